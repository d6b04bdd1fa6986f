// Test helper: stdin -> gfb_text::share_reciprocals -> stdout.
#include <iostream>
#include <iterator>
#include "../../integration/text_passes.hpp"
int main() {
    const std::string in((std::istreambuf_iterator<char> (std::cin)), std::istreambuf_iterator<char> ());
    std::cout << gfb_text::share_reciprocals(in);
    return 0;
}
