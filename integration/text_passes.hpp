//------------------------------------------------------------------------------
//  text_passes.hpp -- JIT passes over kernel TEXT emitted by the reference's own front end, applied by
//  integration/b200_context.hpp before NVRTC sees it.  No dependency on the reference headers, so the
//  passes are unit-tested on their own (tests/cpp/text_pass_case.cpp, tests/test_graph_host.py).
//------------------------------------------------------------------------------
#ifndef gfb_text_passes_hpp
#define gfb_text_passes_hpp

#include <cctype>
#include <cstdlib>
#include <sstream>
#include <string>
#include <unordered_map>

namespace gfb_text {
///  The reference's nodes write one operation per statement (`const double r<id> = a/b;`).  An IEEE
///  FP64 division costs ~10 FP64-pipe instructions plus a guarded slow path on the GPU, and the
///  reference's EFIT kernels divide 336-696 times per step by only ~80 distinct denominators
///  (SURVEY.md App. A.1).  This pass rewrites the kernel TEXT before NVRTC sees it:
///      a/b            ->  a*i<b>   with  const double i<b> = gfb::rcp(b);  once per denominator
///      a/(double)c    ->  a*(1.0/(double)c)
///      pow(a, (double)1.5)  ->  a*sqrt(a)
///      sqrt(a)        ->  gfb::sqrt_from_rsqrt(a, gfb::rsqrt(a))   (refined hardware seed, IEEE at 0 and inf)
///      (x - o)/s  inside a table index  ->  (x - o)*(1.0/s)         (what the full back end does too)
///  gfb::rcp is the hardware seed refined to <= 1 ulp (skeleton.cuh).  The reference compiles its own
///  kernels with -ffast-math (cpu_context.hpp:155-157), i.e. with reciprocal-math, so the arithmetic
///  stays inside what the reference defines.  GFB_B200_IEEE_DIVIDE=1 keeps the text as emitted.
    inline std::string share_reciprocals(const std::string &source) {
        if (std::getenv("GFB_B200_IEEE_DIVIDE")) return source;
        const bool divides_only = std::getenv("GFB_B200_DIVIDES_ONLY") != nullptr;      // A/B switch for measurements
        auto is_register = [] (const std::string &t) {
            if (t.size() < 2 || (t[0] != 'r' && t[0] != 'v' && t[0] != 'o')) return false;
            for (const char ch : t) if (!std::isalnum(static_cast<unsigned char> (ch)) && ch != '_') return false;
            return true;
        };
        std::istringstream in(source);
        std::ostringstream out;
        std::unordered_map<std::string, std::string> inverse;
        std::string line;
        const std::string head = "const double ";
        while (std::getline(in, line)) {
            if (line.find("extern \"C\"") != std::string::npos) inverse.clear();       // registers are per kernel
            const size_t at = line.find(head);
            const size_t eq = line.find(" = ");
            if (line.size() > 400 || at == std::string::npos || eq == std::string::npos || line.back() != ';' ||
                line.find_first_not_of(' ') != at) {
                out << line << '\n';
                continue;
            }
            const std::string indent = line.substr(0, at);
            const std::string name = line.substr(at + head.size(), eq - at - head.size());
            const std::string rhs = line.substr(eq + 3, line.size() - eq - 4);
            const size_t slash = rhs.find('/');
            if (slash != std::string::npos && rhs.find('/', slash + 1) == std::string::npos &&
                rhs.find_first_of(" ,?") == std::string::npos) {
                const std::string num = rhs.substr(0, slash), den = rhs.substr(slash + 1);
                if (is_register(den)) {
                    auto found = inverse.find(den);
                    if (found == inverse.end()) {
                        const std::string iname = "i" + den;
                        out << indent << head << iname << " = gfb::rcp(" << den << ");\n";
                        found = inverse.emplace(den, iname).first;
                    }
                    out << indent << head << name << " = " << num << "*" << found->second << ";\n";
                    continue;
                }
                if (den.rfind("(double)", 0) == 0 && den.find('(', 8) == std::string::npos) {
                    out << indent << head << name << " = " << num << "*(1.0/" << den << ");\n";
                    continue;
                }
            }
            const std::string pow_head = "pow(", pow_tail = ", (double)1.5)";
            if (rhs.rfind(pow_head, 0) == 0 && rhs.size() > pow_head.size() + pow_tail.size() &&
                rhs.compare(rhs.size() - pow_tail.size(), pow_tail.size(), pow_tail) == 0) {
                const std::string base = rhs.substr(pow_head.size(), rhs.size() - pow_head.size() - pow_tail.size());
                if (is_register(base)) {
                    out << indent << head << name << " = " << base << "*gfb::sqrt_from_rsqrt(" << base << ", gfb::rsqrt(" << base << "));\n";
                    continue;
                }
            }
            const std::string sqrt_head = "sqrt(";
            if (!divides_only && rhs.rfind(sqrt_head, 0) == 0 && rhs.back() == ')' &&
                is_register(rhs.substr(sqrt_head.size(), rhs.size() - sqrt_head.size() - 1))) {
                const std::string arg = rhs.substr(sqrt_head.size(), rhs.size() - sqrt_head.size() - 1);
                out << indent << head << name << " = gfb::sqrt_from_rsqrt(" << arg << ", gfb::rsqrt(" << arg << "));\n";
                continue;
            }
//  Table look-ups: `...max<double>((x - offset)/scale,0)...` with a literal scale.
            if (!divides_only && rhs.find("max<double>((") != std::string::npos) {
                std::string edited = rhs;
                size_t pos = 0;
                bool changed = false;
                while ((pos = edited.find(")/", pos)) != std::string::npos) {
                    size_t end = pos + 2;
                    while (end < edited.size() && (std::isdigit(static_cast<unsigned char> (edited[end])) || edited[end] == '.' ||
                                                   edited[end] == 'e' || edited[end] == 'E' || edited[end] == '-' || edited[end] == '+')) end++;
                    if (end > pos + 2 && end < edited.size() && edited[end] == ',') {
                        const std::string with = ")*(1.0/" + edited.substr(pos + 2, end - pos - 2) + ")";
                        edited.replace(pos, end - pos, with);
                        pos += with.size();
                        changed = true;
                    } else {
                        pos += 2;
                    }
                }
                if (changed) {
                    out << indent << head << name << " = " << edited << ";\n";
                    continue;
                }
            }
            out << line << '\n';
        }
        return out.str();
    }
}

#endif /* gfb_text_passes_hpp */
