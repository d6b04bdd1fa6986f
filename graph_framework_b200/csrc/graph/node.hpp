//------------------------------------------------------------------------------
//  node.hpp -- expression graph for the B200 back end.
//
//  Host-side mirror of the reference's graph layer for the FP64 ray path:
//    /root/reference/graph_framework/node.hpp        (leaf_node, constant, variable, pseudo_variable)
//    /root/reference/graph_framework/arithmetic.hpp  (add, subtract, multiply, divide, fma)
//    /root/reference/graph_framework/math.hpp        (sqrt, exp, log, pow)
//    /root/reference/graph_framework/trigonometry.hpp(sin, cos, tan, atan)
//    /root/reference/graph_framework/piecewise.hpp   (piecewise_1D, piecewise_2D)
//  Same public names and meaning (graph::variable, graph::constant, operators,
//  df(), evaluate(), pseudo_variable, remove_pseudo, is_match, variable_cast ...).
//
//  This is NOT a port: one tagged node type instead of a class per operation,
//  every node hash-consed at creation (so common sub-expression elimination is
//  structural and node identity is pointer identity), a small normalising rule
//  set instead of the reference's canonicalising reducer, and code generation
//  that lives in emit.hpp (device-function bodies for hand-written sm_100a
//  skeletons) rather than in per-node compile() methods.
//
//  Only T = double, SAFE_MATH = false is implemented: that is the north-star
//  path (xrays.cpp:1096-1100 runs trace_ray<double>).  The template parameters
//  are kept so reference-style user code compiles unchanged.
//------------------------------------------------------------------------------
#ifndef gfb_graph_node_hpp
#define gfb_graph_node_hpp

#include <array>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../special.cuh"

namespace backend {
//------------------------------------------------------------------------------
///  Host value buffer (reference: backend.hpp:28-786, a std::vector wrapper).
///  A buffer of size one broadcasts against any other size.
//------------------------------------------------------------------------------
    template<typename T=double>
    class buffer {
    private:
        std::vector<T> memory;
    public:
        buffer() {}
        buffer(const size_t s) : memory(s) {}
        buffer(const size_t s, const T d) : memory(s, d) {}
        buffer(const std::vector<T> &d) : memory(d) {}
        T &operator[] (const size_t i) { return memory[i]; }
        const T &operator[] (const size_t i) const { return memory[i]; }
        const T at(const size_t i) const { return memory.at(i); }
        size_t size() const { return memory.size(); }
        T *data() { return memory.data(); }
        const T *data() const { return memory.data(); }
        void set(const T d) { memory.assign(memory.size(), d); }
        void set(const std::vector<T> &d) { memory = d; }
        const std::vector<T> &vec() const { return memory; }
        T max() const {
            T m = memory.at(0);
            for (const T v : memory) m = std::max(m, v);
            return m;
        }
        bool is_same() const {
            for (const T v : memory) if (v != memory[0]) return false;
            return true;
        }
    };
}

namespace jit {
    template<typename T>
    concept float_scalar = std::is_same<T, double>::value;
}

namespace graph {
    enum class op_t : uint8_t {
        constant, variable, pseudo,
        add, sub, mul, div, fma,
        sqrt, exp, log, pow,
        sin, cos, atan,
        piecewise_1d, piecewise_2d,
        fourier,
        erfi, nan_to_zero,
        index_1d, index_2d,
        spline_1d, spline_2d
    };

    class leaf_node;
    using leaf_ptr = std::shared_ptr<leaf_node>;

///  Convenience type alias matching graph::shared_leaf<T, SAFE_MATH> (node.hpp:676).
    template<typename T=double, bool SAFE_MATH=false>
    using shared_leaf = leaf_ptr;
    template<typename T=double, bool SAFE_MATH=false>
    using shared_variable = leaf_ptr;
    template<typename T=double, bool SAFE_MATH=false>
    using output_nodes = std::vector<leaf_ptr>;
    template<typename T=double, bool SAFE_MATH=false>
    using input_nodes = std::vector<leaf_ptr>;
    template<typename T=double, bool SAFE_MATH=false>
    using map_nodes = std::vector<std::pair<leaf_ptr, leaf_ptr>>;
///  Random states are not on the FP64 ray path (random.hpp); the slot is kept
///  so add_item(...) call sites keep their shape.  It must be empty.
    struct random_state_placeholder {
        void *get() const { return nullptr; }
    };
    template<typename T=double, bool SAFE_MATH=false>
    using shared_random_state = random_state_placeholder;

//------------------------------------------------------------------------------
///  Coefficient table shared by piecewise nodes.  Tables are interned by
///  content so identical data gives identical nodes.
//------------------------------------------------------------------------------
    struct table_data {
        std::vector<double> values;
        uint64_t hash;
///  Fourier-series tables only (op_t::fourier): values = folded cubic coefficients
///  [mode][cell][4], xm/xn = poloidal/toroidal mode numbers.
        std::vector<double> xm, xn;
        size_t cells = 0;
    };
    using table_ptr = std::shared_ptr<const table_data>;

//------------------------------------------------------------------------------
///  The one node type.
//------------------------------------------------------------------------------
    class leaf_node : public std::enable_shared_from_this<leaf_node> {
    public:
        const op_t op;
        const std::array<leaf_ptr, 3> args;
///  Constant value (op_t::constant).
        const double value;
///  Piecewise table, dimensions and argument normalisation
///  index = trunc(clamp((arg - offset)/scale, 0, n - 1)), piecewise.hpp:26-65.
        const table_ptr table;
        const size_t num_cols;
        const std::array<double, 2> scale;
        const std::array<double, 2> offset;
///  Creation order, used for canonical argument order and register names.
        const uint64_t id;
///  Variable storage (op_t::variable) and symbol.
        std::vector<double> buffer;
        std::string symbol;

        leaf_node(const op_t op, std::array<leaf_ptr, 3> a, const double v,
                  table_ptr t, const size_t nc,
                  std::array<double, 2> s, std::array<double, 2> o,
                  const uint64_t id) :
        op(op), args(a), value(v), table(t), num_cols(nc), scale(s), offset(o), id(id) {}

        size_t num_args() const {
            switch (op) {
                case op_t::constant: case op_t::variable: return 0;
                case op_t::pseudo: case op_t::sqrt: case op_t::exp: case op_t::log:
                case op_t::sin: case op_t::cos: case op_t::piecewise_1d: case op_t::erfi:
                case op_t::nan_to_zero: case op_t::spline_1d: return 1;
                case op_t::fma: case op_t::fourier: case op_t::index_2d: return 3;
                default: return 2;
            }
        }

        bool is_constant() const { return op == op_t::constant; }
        bool is_constant(const double v) const { return op == op_t::constant && value == v; }
        bool is_piecewise() const { return op == op_t::piecewise_1d || op == op_t::piecewise_2d; }
        bool is_index() const { return op == op_t::index_1d || op == op_t::index_2d; }
        bool is_spline() const { return op == op_t::spline_1d || op == op_t::spline_2d; }

//  -- reference API -----------------------------------------------------------
///  Host evaluation (node.hpp:378 evaluate()).
        backend::buffer<double> evaluate();
///  Symbolic derivative with respect to any node (node.hpp:396 df()).
        leaf_ptr df(leaf_ptr x);
///  Nodes are reduced when they are built; kept for API compatibility.
        leaf_ptr reduce() { return shared_from_this(); }
///  Structural match == pointer identity because every node is interned.
        bool is_match(leaf_ptr x) { return x.get() == this; }
///  Replace pseudo variables by the expressions they wrap (node.hpp:1745).
        leaf_ptr remove_pseudo();
///  True when the value does not depend on any variable.
        bool is_constant_like();

//  -- variable API (node.hpp:1386 variable_node) --------------------------------
        size_t size() const { return buffer.size(); }
        void set(const double d) { assert(op == op_t::variable); buffer.assign(buffer.size(), d); }
        void set(const size_t index, const double d) { assert(op == op_t::variable); buffer.at(index) = d; }
        void set(const std::vector<double> &d) { assert(op == op_t::variable); buffer = d; }
        void set(const backend::buffer<double> &d) { assert(op == op_t::variable); buffer = d.vec(); }
        double *data() { return buffer.data(); }
        const std::string &get_symbol() const { return symbol; }

        std::string to_string();
        void to_latex() { std::cout << to_string(); }
    };

//------------------------------------------------------------------------------
//  Interning.
//------------------------------------------------------------------------------
    namespace detail {
        struct key {
            op_t op;
            const leaf_node *a, *b, *c;
            uint64_t bits;
            const table_data *table;
            uint64_t nc;
            uint64_t s0, s1, o0, o1;
            bool operator==(const key &k) const {
                return op == k.op && a == k.a && b == k.b && c == k.c && bits == k.bits &&
                       table == k.table && nc == k.nc && s0 == k.s0 && s1 == k.s1 &&
                       o0 == k.o0 && o1 == k.o1;
            }
        };
        struct key_hash {
            size_t operator()(const key &k) const {
                uint64_t h = 1469598103934665603ull;
                auto mix = [&h] (const uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
                mix(static_cast<uint64_t> (k.op));
                mix(reinterpret_cast<uint64_t> (k.a)); mix(reinterpret_cast<uint64_t> (k.b));
                mix(reinterpret_cast<uint64_t> (k.c)); mix(k.bits);
                mix(reinterpret_cast<uint64_t> (k.table)); mix(k.nc);
                mix(k.s0); mix(k.s1); mix(k.o0); mix(k.o1);
                return static_cast<size_t> (h);
            }
        };
        struct pair_hash {
            size_t operator()(const std::pair<const leaf_node *, const leaf_node *> &p) const {
                return std::hash<const void *> () (p.first)*1000003u ^ std::hash<const void *> () (p.second);
            }
        };
///  Per-thread caches, as in the reference (node.hpp:660-668): node pointers are
///  thread private, one graph per device thread.
        struct caches_t {
            std::unordered_map<key, leaf_ptr, key_hash> nodes;
            std::unordered_map<uint64_t, std::vector<table_ptr>> tables;
            std::unordered_map<std::pair<const leaf_node *, const leaf_node *>, leaf_ptr, pair_hash> df;
            std::unordered_map<const leaf_node *, leaf_ptr> no_pseudo;
            uint64_t next_id = 0;
            bool fold_tables = false;
        };
        inline caches_t &caches() {
            static thread_local caches_t c;
            return c;
        }
        inline uint64_t bits(const double d) {
            uint64_t u;
            std::memcpy(&u, &d, sizeof(u));
            return u;
        }
        inline leaf_ptr intern(const op_t op, leaf_ptr a, leaf_ptr b, leaf_ptr c,
                               const double v=0.0, table_ptr t=table_ptr(), const size_t nc=0,
                               std::array<double, 2> s={0.0, 0.0}, std::array<double, 2> o={0.0, 0.0}) {
            auto &cc = caches();
            const key k = {op, a.get(), b.get(), c.get(), bits(v), t.get(), nc,
                           bits(s[0]), bits(s[1]), bits(o[0]), bits(o[1])};
            auto it = cc.nodes.find(k);
            if (it != cc.nodes.end()) {
                return it->second;
            }
            auto n = std::make_shared<leaf_node> (op, std::array<leaf_ptr, 3> {a, b, c}, v, t, nc, s, o, cc.next_id++);
            cc.nodes.emplace(k, n);
            return n;
        }
        inline table_ptr intern_table(const std::vector<double> &values) {
            uint64_t h = 1469598103934665603ull;
            for (const double v : values) {
                h ^= bits(v);
                h *= 1099511628211ull;
            }
            auto &bucket = caches().tables[h];
            for (auto &t : bucket) {
                if (t->values.size() == values.size() &&
                    !std::memcmp(t->values.data(), values.data(), sizeof(double)*values.size())) {
                    return t;
                }
            }
            auto t = std::make_shared<table_data> ();
            t->values = values;
            t->hash = h;
            bucket.push_back(t);
            return t;
        }
    }

///  Drop every cached node of this thread (nodes held by the user stay valid).
    inline void clear_caches() {
        auto &c = detail::caches();
        c.nodes.clear();
        c.tables.clear();
        c.df.clear();
        c.no_pseudo.clear();
    }

//------------------------------------------------------------------------------
//  Leaves.
//------------------------------------------------------------------------------
    template<typename T=double, bool SAFE_MATH=false>
    shared_leaf<T, SAFE_MATH> constant(const double d) {
        static_assert(std::is_same<T, double>::value && !SAFE_MATH, "B200 back end: double, SAFE_MATH=false only.");
        return detail::intern(op_t::constant, nullptr, nullptr, nullptr, d);
    }
    template<typename T=double, bool SAFE_MATH=false>
    shared_leaf<T, SAFE_MATH> zero() { return constant<T, SAFE_MATH> (0.0); }
    template<typename T=double, bool SAFE_MATH=false>
    shared_leaf<T, SAFE_MATH> one() { return constant<T, SAFE_MATH> (1.0); }
    template<typename T=double, bool SAFE_MATH=false>
    shared_leaf<T, SAFE_MATH> none() { return constant<T, SAFE_MATH> (-1.0); }

///  Variables are never interned: each one is a distinct buffer.
    template<typename T=double, bool SAFE_MATH=false>
    shared_leaf<T, SAFE_MATH> variable(const size_t size, const std::string &symbol) {
        static_assert(std::is_same<T, double>::value && !SAFE_MATH, "B200 back end: double, SAFE_MATH=false only.");
        auto n = std::make_shared<leaf_node> (op_t::variable, std::array<leaf_ptr, 3> {}, 0.0, table_ptr(), 0,
                                              std::array<double, 2> {0.0, 0.0}, std::array<double, 2> {0.0, 0.0},
                                              detail::caches().next_id++);
        n->buffer.assign(size, 0.0);
        n->symbol = symbol;
        return n;
    }
    template<typename T=double, bool SAFE_MATH=false>
    shared_leaf<T, SAFE_MATH> variable(const size_t size, const double d, const std::string &symbol) {
        auto n = variable<T, SAFE_MATH> (size, symbol);
        n->set(d);
        return n;
    }
    template<typename T=double, bool SAFE_MATH=false>
    shared_leaf<T, SAFE_MATH> variable(const std::vector<double> &d, const std::string &symbol) {
        auto n = variable<T, SAFE_MATH> (d.size(), symbol);
        n->set(d);
        return n;
    }

///  Cast helpers (node.hpp variable_cast / constant_cast / pseudo_variable_cast):
///  return the node when it has that kind, an empty pointer otherwise.
    inline leaf_ptr variable_cast(leaf_ptr x) { return x.get() && x->op == op_t::variable ? x : leaf_ptr(); }
    inline leaf_ptr constant_cast(leaf_ptr x) { return x.get() && x->op == op_t::constant ? x : leaf_ptr(); }
    inline leaf_ptr pseudo_variable_cast(leaf_ptr x) { return x.get() && x->op == op_t::pseudo ? x : leaf_ptr(); }
    inline leaf_ptr piecewise_1D_cast(leaf_ptr x) { return x.get() && x->op == op_t::piecewise_1d ? x : leaf_ptr(); }
    inline leaf_ptr piecewise_2D_cast(leaf_ptr x) { return x.get() && x->op == op_t::piecewise_2d ? x : leaf_ptr(); }

///  A pseudo variable hides its argument from df() (node.hpp:1745): derivatives
///  treat it as an independent leaf; evaluation and code generation see through it.
    template<typename T=double, bool SAFE_MATH=false>
    shared_leaf<T, SAFE_MATH> pseudo_variable(shared_leaf<T, SAFE_MATH> x) {
        return detail::intern(op_t::pseudo, x, nullptr, nullptr);
    }

//------------------------------------------------------------------------------
//  Piecewise helpers.
//------------------------------------------------------------------------------
    namespace detail {
        inline bool same_cells(const leaf_node *a, const leaf_node *b) {
            return a->op == b->op && a->args[0] == b->args[0] && a->args[1] == b->args[1] &&
                   a->num_cols == b->num_cols && a->scale == b->scale && a->offset == b->offset &&
                   a->table->values.size() == b->table->values.size();
        }
        inline leaf_ptr with_table(const leaf_node *like, const std::vector<double> &values) {
            return intern(like->op, like->args[0], like->args[1], nullptr, 0.0, intern_table(values),
                          like->num_cols, like->scale, like->offset);
        }
        template<typename F>
        inline leaf_ptr map_table(const leaf_node *p, F f) {
            std::vector<double> v(p->table->values);
            for (double &e : v) e = f(e);
            return with_table(p, v);
        }
        template<typename F>
        inline leaf_ptr zip_table(const leaf_node *a, const leaf_node *b, F f) {
            std::vector<double> v(a->table->values);
            for (size_t i = 0; i < v.size(); i++) v[i] = f(v[i], b->table->values[i]);
            return with_table(a, v);
        }
    }

//------------------------------------------------------------------------------
///  RAII switch: while alive, arithmetic between piecewise nodes on the same
///  cells and constants is folded into new tables on the host, the way the
///  reference's reducer does it everywhere (piecewise.hpp:120-236).  Used only
///  while spline coefficients are built (equilibrium.hpp build_1D_spline) so
///  derivative algebra does not multiply the number of tables.
//------------------------------------------------------------------------------
    struct fold_tables_scope {
        const bool previous;
        fold_tables_scope(const bool enable=true) : previous(detail::caches().fold_tables) {
            detail::caches().fold_tables = enable;
        }
        ~fold_tables_scope() { detail::caches().fold_tables = previous; }
    };

//------------------------------------------------------------------------------
//  Arithmetic.  Normalisation happens here, at construction.
//------------------------------------------------------------------------------
    inline leaf_ptr add(leaf_ptr l, leaf_ptr r);
    inline leaf_ptr sub(leaf_ptr l, leaf_ptr r);
    inline leaf_ptr mul(leaf_ptr l, leaf_ptr r);
    inline leaf_ptr div(leaf_ptr l, leaf_ptr r);
    inline leaf_ptr fma(leaf_ptr a, leaf_ptr b, leaf_ptr c);
    inline leaf_ptr sqrt(leaf_ptr a);
    inline leaf_ptr exp(leaf_ptr a);
    inline leaf_ptr log(leaf_ptr a);
    inline leaf_ptr erfi(leaf_ptr a);
    inline leaf_ptr nan_to_zero(leaf_ptr a);
    inline leaf_ptr pow(leaf_ptr a, leaf_ptr b);
    inline leaf_ptr sin(leaf_ptr a);
    inline leaf_ptr cos(leaf_ptr a);
    inline leaf_ptr tan(leaf_ptr a);
    inline leaf_ptr atan(leaf_ptr x, leaf_ptr y);

    inline leaf_ptr add(leaf_ptr l, leaf_ptr r) {
        if (l->is_constant() && r->is_constant()) return constant(l->value + r->value);
        if (l->is_constant(0.0)) return r;
        if (r->is_constant(0.0)) return l;
        if (detail::caches().fold_tables) {
            if (l->is_piecewise() && r->is_piecewise() && detail::same_cells(l.get(), r.get()))
                return detail::zip_table(l.get(), r.get(), [] (double a, double b) { return a + b; });
            if (l->is_piecewise() && r->is_constant())
                return detail::map_table(l.get(), [&r] (double a) { return a + r->value; });
            if (r->is_piecewise() && l->is_constant())
                return detail::map_table(r.get(), [&l] (double a) { return l->value + a; });
        }
        if (l == r) return mul(constant(2.0), l);
        if (l->id > r->id) std::swap(l, r);
        return detail::intern(op_t::add, l, r, nullptr);
    }

    inline leaf_ptr sub(leaf_ptr l, leaf_ptr r) {
        if (l->is_constant() && r->is_constant()) return constant(l->value - r->value);
        if (r->is_constant(0.0)) return l;
        if (l == r) return zero();
        if (detail::caches().fold_tables) {
            if (l->is_piecewise() && r->is_piecewise() && detail::same_cells(l.get(), r.get()))
                return detail::zip_table(l.get(), r.get(), [] (double a, double b) { return a - b; });
            if (l->is_piecewise() && r->is_constant())
                return detail::map_table(l.get(), [&r] (double a) { return a - r->value; });
            if (r->is_piecewise() && l->is_constant())
                return detail::map_table(r.get(), [&l] (double a) { return l->value - a; });
        }
        if (l->is_constant(0.0)) return mul(none(), r);
        return detail::intern(op_t::sub, l, r, nullptr);
    }

    inline leaf_ptr mul(leaf_ptr l, leaf_ptr r) {
        if (l->is_constant() && r->is_constant()) return constant(l->value*r->value);
        if (l->is_constant(0.0) || r->is_constant(0.0)) return zero();
        if (l->is_constant(1.0)) return r;
        if (r->is_constant(1.0)) return l;
        if (r->is_constant()) std::swap(l, r);
        if (detail::caches().fold_tables && l->is_constant() && r->is_piecewise())
            return detail::map_table(r.get(), [&l] (double a) { return l->value*a; });
        if (detail::caches().fold_tables && l->is_piecewise() && r->is_piecewise() &&
            detail::same_cells(l.get(), r.get()))
            return detail::zip_table(l.get(), r.get(), [] (double a, double b) { return a*b; });
        if (l == r) {
//  sqrt(u)*sqrt(u) -> u and (p*sqrt(u))*(p*sqrt(u)) -> (p*p)*u.  Besides saving the square root this
//  keeps sqrt out of expressions that only need its square (n_perp^2 and (n_par n_perp)^2 in the
//  cold-plasma determinant, dispersion.hpp:990-1007), whose derivative would otherwise be 0/0
//  when k is parallel to B -- the state of the reference's cut-off searches (physics_test.cpp:472-530).
            if (l->op == op_t::sqrt) return l->args[0];
            if (l->op == op_t::mul && l->args[1]->op == op_t::sqrt)
                return mul(mul(l->args[0], l->args[0]), l->args[1]->args[0]);
            if (l->op == op_t::mul && l->args[0]->op == op_t::sqrt)
                return mul(mul(l->args[1], l->args[1]), l->args[0]->args[0]);
        }
        if (l->is_constant()) {
//  c1*(c2*x) -> (c1*c2)*x
            if (r->op == op_t::mul && r->args[0]->is_constant())
                return mul(constant(l->value*r->args[0]->value), r->args[1]);
//  c1*(c2/x) -> (c1*c2)/x
            if (r->op == op_t::div && r->args[0]->is_constant())
                return div(constant(l->value*r->args[0]->value), r->args[1]);
        } else if (l->id > r->id) {
            std::swap(l, r);
        }
        return detail::intern(op_t::mul, l, r, nullptr);
    }

    inline leaf_ptr div(leaf_ptr l, leaf_ptr r) {
        if (l->is_constant() && r->is_constant()) return constant(l->value/r->value);
        if (l->is_constant(0.0)) return zero();
        if (r->is_constant(1.0)) return l;
        if (l == r) return one();
        if (detail::caches().fold_tables) {
            if (l->is_piecewise() && r->is_constant())
                return detail::map_table(l.get(), [&r] (double a) { return a/r->value; });
            if (l->is_piecewise() && r->is_piecewise() && detail::same_cells(l.get(), r.get()))
                return detail::zip_table(l.get(), r.get(), [] (double a, double b) { return a/b; });
        }
//  x/c -> (1/c)*x : one reciprocal at graph-build time instead of a divide per ray.
        if (r->is_constant()) return mul(constant(1.0/r->value), l);
//  (c*x)/y -> c*(x/y) keeps constants outermost where they merge.
        if (l->op == op_t::mul && l->args[0]->is_constant())
            return mul(l->args[0], div(l->args[1], r));
//  x/(c*y) -> (1/c)*(x/y): the denominator is then y itself, whose reciprocal the kernel often has already
//  (d sqrt(u) = du/(2 sqrt(u)) divides by the root, and 1/sqrt(u) comes for free with it).
        if (r->op == op_t::mul && r->args[0]->is_constant())
            return mul(constant(1.0/r->args[0]->value), div(l, r->args[1]));
        return detail::intern(op_t::div, l, r, nullptr);
    }

    inline leaf_ptr fma(leaf_ptr a, leaf_ptr b, leaf_ptr c) {
        if (a->is_constant() && b->is_constant() && c->is_constant())
            return constant(std::fma(a->value, b->value, c->value));
        if (a->is_constant(0.0) || b->is_constant(0.0)) return c;
        if (c->is_constant(0.0)) return mul(a, b);
        if (a->is_constant(1.0)) return add(b, c);
        if (b->is_constant(1.0)) return add(a, c);
        if (a->is_constant() && b->is_constant()) return add(constant(a->value*b->value), c);
        if (b->is_constant() || (!a->is_constant() && a->id > b->id)) std::swap(a, b);
        return detail::intern(op_t::fma, a, b, c);
    }

    inline leaf_ptr sqrt(leaf_ptr a) {
        if (a->is_constant()) return constant(std::sqrt(a->value));
        return detail::intern(op_t::sqrt, a, nullptr, nullptr);
    }
    inline leaf_ptr exp(leaf_ptr a) {
        if (a->is_constant()) return constant(std::exp(a->value));
        return detail::intern(op_t::exp, a, nullptr, nullptr);
    }
    inline leaf_ptr log(leaf_ptr a) {
        if (a->is_constant()) return constant(std::log(a->value));
        return detail::intern(op_t::log, a, nullptr, nullptr);
    }
///  erfi of a real argument (reference: graph::erfi, math.hpp; evaluation special_functions.hpp:1583).
    inline leaf_ptr erfi(leaf_ptr a) {
        if (a->is_constant()) return constant(gfb::erfi(a->value));
        return detail::intern(op_t::erfi, a, nullptr, nullptr);
    }
///  x if x == x, else 0: the store rule of the reference's SAFE_MATH kernels
///  (cpu_context.hpp:533-544, cuda_context.hpp:905-930) made available as a node.
    inline leaf_ptr nan_to_zero(leaf_ptr a) {
        if (a->is_constant()) return constant(a->value == a->value ? a->value : 0.0);
        if (a->op == op_t::nan_to_zero) return a;
        return detail::intern(op_t::nan_to_zero, a, nullptr, nullptr);
    }
    inline leaf_ptr pow(leaf_ptr a, leaf_ptr b) {
        if (a->is_constant() && b->is_constant()) return constant(std::pow(a->value, b->value));
        if (b->is_constant()) {
            const double e = b->value;
            if (e == 0.0) return one();
            if (e == 1.0) return a;
            if (e == 0.5) return sqrt(a);
            if (e == -0.5) return div(one(), sqrt(a));
            if (e == 1.5) return mul(a, sqrt(a));
            if (e == std::floor(e) && std::abs(e) <= 16.0) {
//  Integer powers are unrolled into multiplies (math.hpp:1215-1227).
                leaf_ptr result = a;
                for (int i = 1, ie = static_cast<int> (std::abs(e)); i < ie; i++) result = mul(result, a);
                return e > 0.0 ? result : div(one(), result);
            }
        }
        return detail::intern(op_t::pow, a, b, nullptr);
    }
    inline leaf_ptr sin(leaf_ptr a) {
        if (a->is_constant()) return constant(std::sin(a->value));
//  sin(atan(x, y)) -> y/sqrt(x^2 + y^2)  (trigonometry.hpp:85-91)
        if (a->op == op_t::atan) {
            auto x = a->args[0], y = a->args[1];
            return div(y, sqrt(add(mul(x, x), mul(y, y))));
        }
        return detail::intern(op_t::sin, a, nullptr, nullptr);
    }
    inline leaf_ptr cos(leaf_ptr a) {
        if (a->is_constant()) return constant(std::cos(a->value));
//  cos(atan(x, y)) -> x/sqrt(x^2 + y^2)  (trigonometry.hpp:342-348)
        if (a->op == op_t::atan) {
            auto x = a->args[0], y = a->args[1];
            return div(x, sqrt(add(mul(x, x), mul(y, y))));
        }
        return detail::intern(op_t::cos, a, nullptr, nullptr);
    }
    inline leaf_ptr tan(leaf_ptr a) { return div(sin(a), cos(a)); }
///  atan(x, y) is the angle of the point (x, y): atan2(y, x) (trigonometry.hpp:711-722).
    inline leaf_ptr atan(leaf_ptr x, leaf_ptr y) {
        if (x->is_constant() && y->is_constant()) return constant(std::atan2(y->value, x->value));
        return detail::intern(op_t::atan, x, y, nullptr);
    }

//------------------------------------------------------------------------------
///  piecewise_1D: coefficient looked up by a clamped, truncated normalised
///  argument (piecewise.hpp:256-325); df() of a coefficient is zero (:241-243).
//------------------------------------------------------------------------------
    template<typename T=double, bool SAFE_MATH=false>
    shared_leaf<T, SAFE_MATH> piecewise_1D(const backend::buffer<T> &d, shared_leaf<T, SAFE_MATH> x,
                                           const T scale, const T offset) {
        if (d.is_same()) return constant(d.at(0));
        return detail::intern(op_t::piecewise_1d, x, nullptr, nullptr, 0.0, detail::intern_table(d.vec()), 0,
                              {scale, 0.0}, {offset, 0.0});
    }
    template<typename T=double, bool SAFE_MATH=false>
    shared_leaf<T, SAFE_MATH> piecewise_1D(const std::vector<T> &d, shared_leaf<T, SAFE_MATH> x,
                                           const T scale, const T offset) {
        return piecewise_1D(backend::buffer<T> (d), x, scale, offset);
    }
///  piecewise_2D: row-major table, index = i_x*num_cols + i_y (piecewise.hpp:1195-1201).
    template<typename T=double, bool SAFE_MATH=false>
    shared_leaf<T, SAFE_MATH> piecewise_2D(const backend::buffer<T> &d, const size_t num_cols,
                                           shared_leaf<T, SAFE_MATH> x, const T x_scale, const T x_offset,
                                           shared_leaf<T, SAFE_MATH> y, const T y_scale, const T y_offset) {
        assert(d.size()%num_cols == 0 && "Table size must be a multiple of the number of columns.");
        if (d.is_same()) return constant(d.at(0));
        return detail::intern(op_t::piecewise_2d, x, y, nullptr, 0.0, detail::intern_table(d.vec()), num_cols,
                              {x_scale, y_scale}, {x_offset, y_offset});
    }
    template<typename T=double, bool SAFE_MATH=false>
    shared_leaf<T, SAFE_MATH> piecewise_2D(const std::vector<T> &d, const size_t num_cols,
                                           shared_leaf<T, SAFE_MATH> x, const T x_scale, const T x_offset,
                                           shared_leaf<T, SAFE_MATH> y, const T y_scale, const T y_offset) {
        return piecewise_2D(backend::buffer<T> (d), num_cols, x, x_scale, x_offset, y, y_scale, y_offset);
    }

//------------------------------------------------------------------------------
///  index_1D / index_2D: gather from a VARIABLE (a device array that kernels may rewrite between
///  launches) with the piecewise index rule (piecewise.hpp:1436-1640, 1776-2010):
///      index_1D(v, x) = v[trunc(clamp((x - offset)/scale, 0, size - 1))]
///      index_2D(v, x, y) = v[i_x*num_cols + i_y].
///  The indexed variable may have any length; a kernel that only indexes it never loads it per ray.
///  df() is 1 with respect to the node itself and 0 otherwise (piecewise.hpp:1475-1477).
//------------------------------------------------------------------------------
    template<typename T=double, bool SAFE_MATH=false>
    shared_leaf<T, SAFE_MATH> index_1D(shared_leaf<T, SAFE_MATH> variable, shared_leaf<T, SAFE_MATH> x,
                                       const T scale, const T offset) {
        assert(variable_cast(variable).get() && "index_1D needs a variable to index.");
        return detail::intern(op_t::index_1d, variable, x, nullptr, 0.0, table_ptr(), 0, {scale, 0.0}, {offset, 0.0});
    }
    template<typename T=double, bool SAFE_MATH=false>
    shared_leaf<T, SAFE_MATH> index_2D(shared_leaf<T, SAFE_MATH> variable, const size_t num_cols,
                                       shared_leaf<T, SAFE_MATH> x, const T x_scale, const T x_offset,
                                       shared_leaf<T, SAFE_MATH> y, const T y_scale, const T y_offset) {
        assert(variable_cast(variable).get() && "index_2D needs a variable to index.");
        assert(variable->size()%num_cols == 0 && "Variable size must be a multiple of the number of columns.");
        return detail::intern(op_t::index_2d, variable, x, y, 0.0, table_ptr(), num_cols,
                              {x_scale, y_scale}, {x_offset, y_offset});
    }

//------------------------------------------------------------------------------
///  Fourier series with radial spline amplitudes -- the building block of the VMEC equilibrium
///  (equilibrium.hpp:2120-2151 sums 86 modes of  spline_mn(s) * {cos, sin}(m u - n v)).
///
///      F[a, b, c](s, u, v) = sum_mn  d^a P_mn/ds^a (s) * d^(b+c)/du^b dv^c  trig(m u - n v)
///
///  P_mn is the cell's cubic in the physical s (coefficients folded like build_1D_spline,
///  equilibrium.hpp:1121-1133; constant within a cell, so their derivative is zero like a
///  piecewise node's).  The family is closed under differentiation, so df() and gradient() stay
///  inside it, and the emitter evaluates every member that shares (u, v, mode numbers) in ONE
///  device loop over the modes instead of unrolling ~10^4 statements.
///  num_cols packs (a, b, c, base): base 0 = cos series, 1 = sin series.
//------------------------------------------------------------------------------
    inline size_t table_index(const double x, const double scale, const double offset, const size_t n);
    struct fourier_order {
        unsigned a, b, c, base;
        size_t pack() const { return a | (b << 8) | (c << 16) | (base << 24); }
        static fourier_order unpack(const size_t p) {
            return {static_cast<unsigned> (p & 255), static_cast<unsigned> ((p >> 8) & 255),
                    static_cast<unsigned> ((p >> 16) & 255), static_cast<unsigned> ((p >> 24) & 255)};
        }
    };

///  Intern a coefficient set.  c[k][mode][cell] are the reference's raw spline tables c0..c3;
///  they are folded to physical-s coefficients here, in build_1D_spline's operation order.
    inline table_ptr fourier_table(const std::array<std::vector<std::vector<double>>, 4> &c,
                                   const std::vector<double> &xm, const std::vector<double> &xn,
                                   const double scale, const double offset) {
        const size_t modes = xm.size(), cells = c[0].at(0).size();
        std::vector<double> v(modes*cells*4);
        const double s2 = scale*scale, s3 = scale*scale*scale;
        for (size_t m = 0; m < modes; m++) {
            for (size_t i = 0; i < cells; i++) {
                const double c0 = c[0][m][i], c1 = c[1][m][i], c2 = c[2][m][i], c3 = c[3][m][i];
                double *o = &v[(m*cells + i)*4];
                o[3] = c3/s3;
                o[2] = c2/s2 - (3.0*offset)*c3/s3;
                o[1] = c1/scale - (2.0*offset)*c2/s2 + (3.0*offset*offset)*c3/s3;
                o[0] = c0 - offset*c1/scale + (offset*offset)*c2/s2 - (offset*offset*offset)*c3/s3;
            }
        }
        for (const double e : xm) v.push_back(e);      // mode numbers take part in the content hash
        for (const double e : xn) v.push_back(e);
        auto base = detail::intern_table(v);
        if (base->cells == 0) {
            auto t = std::const_pointer_cast<table_data> (base);
            t->xm = xm;
            t->xn = xn;
            t->cells = cells;
        }
        return base;
    }

    inline leaf_ptr fourier_series(table_ptr table, leaf_ptr s, leaf_ptr u, leaf_ptr v,
                                   const double scale, const double offset, const fourier_order order) {
        if (order.a > 3) return zero();
        return detail::intern(op_t::fourier, s, u, v, 0.0, table, order.pack(), {scale, 0.0}, {offset, 0.0});
    }

///  Host value of one member (also the oracle of the device loop in tests).
    inline double fourier_value(const table_data &t, const fourier_order o, const double scale, const double offset,
                                const double s, const double u, const double v) {
        const size_t modes = t.xm.size();
        const size_t cell = table_index(s, scale, offset, t.cells);
        double sum = 0.0;
        for (size_t m = 0; m < modes; m++) {
            const double *c = &t.values[(m*t.cells + cell)*4];
            double p;
            switch (o.a) {
                case 0: p = std::fma(std::fma(std::fma(c[3], s, c[2]), s, c[1]), s, c[0]); break;
                case 1: p = std::fma(std::fma(3.0*c[3], s, 2.0*c[2]), s, c[1]); break;
                case 2: p = std::fma(6.0*c[3], s, 2.0*c[2]); break;
                default: p = 6.0*c[3]; break;
            }
            const double angle = t.xm[m]*u - t.xn[m]*v;
            const unsigned k = (o.b + o.c + (o.base ? 3u : 0u)) & 3u;      // cos, -sin, -cos, sin; sin = cos shifted by 3
            const double trig = k == 0 ? std::cos(angle) : k == 1 ? -std::sin(angle) : k == 2 ? -std::cos(angle) : std::sin(angle);
            double factor = 1.0;
            for (unsigned i = 0; i < o.b; i++) factor *= t.xm[m];
            for (unsigned i = 0; i < o.c; i++) factor *= -t.xn[m];
            sum += p*factor*trig;
        }
        return sum;
    }

//------------------------------------------------------------------------------
///  Cubic splines as ONE node per evaluated quantity -- the EFIT equilibrium's building blocks
///  (equilibrium.hpp:1121-1133 build_1D_spline, :1279-1313 build_psi).
///
///      G[a](x)       = d^a/dx^a      sum_j   c_j  x^j                    (cell picked by x)
///      F[a, b](r, z) = d^(a+b)/dr^a dz^b  sum_ij c_ij u^i z^j,  u = (r - offset_r)/scale_r  (cell by r, z)
///
///  with the cell's coefficients folded to the physical argument exactly as build_1D_spline does
///  (F: folded in z, normalised in r, like build_psi).  Coefficients are constant within a cell, so --
///  like a piecewise node's -- their derivative is zero and each family is closed under df() and
///  gradient(): the ray equations need psi, its two first and three second derivatives, and the emitter
///  evaluates all six in one pass over the cell's 16 coefficients (38 FMAs) instead of differentiating
///  through the Horner chains (the reference ends up with 90 folded tables; reverse mode through the
///  chains costs ~3x the direct form).  num_cols packs (a, b, columns of the 2-D grid).
///
///  Evaluation order, shared by host and device so they agree to the last bit except for FMA
///  contraction: value and first derivative follow the Horner chain and its derivative recurrence
///  (the forms df() gives the reference), the second derivative extends the recurrence once more:
///      p1 = x k3 + k2, p2 = x p1 + k1, p3 = x p2 + k0          value
///      q1 = x k3 + p1, q2 = x q1 + p2                          first derivative
///      s1 = x k3 + q1                                          second derivative = 2 s1, third = 6 k3
//------------------------------------------------------------------------------
    struct spline_order {
        unsigned a, b;
        size_t columns;
        size_t pack() const { return a | (b << 8) | (columns << 16); }
        static spline_order unpack(const size_t p) {
            return {static_cast<unsigned> (p & 255), static_cast<unsigned> ((p >> 8) & 255), p >> 16};
        }
    };
    namespace detail {
///  d^a/dx^a of the cubic with coefficients k[0..3] (powers 0..3).
        inline double cubic(const double *k, const double x, const unsigned a) {
            const double p1 = std::fma(x, k[3], k[2]);
            if (a == 3) return 6.0*k[3];
            const double q1 = std::fma(x, k[3], p1);
            if (a == 2) return 2.0*std::fma(x, k[3], q1);
            const double p2 = std::fma(x, p1, k[1]);
            if (a == 1) return std::fma(x, q1, p2);
            return std::fma(x, p2, k[0]);
        }
///  Values of a table-like node (piecewise on `cells` cells, or a constant) as one vector.
        inline std::vector<double> cell_values(const leaf_ptr &n, const size_t cells) {
            if (n->is_constant()) return std::vector<double> (cells, n->value);
            assert(n->is_piecewise() && n->table->values.size() == cells && "Spline coefficients must be tables on the same cells.");
            return n->table->values;
        }
    }

///  c[j]: coefficient of x^j, piecewise_1D nodes (or constants) on the same cells, already folded.
    inline leaf_ptr spline_1d(const std::array<leaf_ptr, 4> &c, leaf_ptr x, const unsigned a=0) {
        const leaf_node *shape = nullptr;
        for (auto &n : c) if (n->op == op_t::piecewise_1d) shape = n.get();
        if (!shape) {           // all coefficients constant: a plain cubic, differentiated by df() like any expression
            leaf_ptr cubic = fma(fma(fma(c[3], x, c[2]), x, c[1]), x, c[0]);
            for (unsigned i = 0; i < a; i++) cubic = cubic->df(x);
            return cubic;
        }
        const size_t cells = shape->table->values.size();
        std::vector<double> v(cells*4);
        for (size_t j = 0; j < 4; j++) {
            const auto col = detail::cell_values(c[j], cells);
            for (size_t i = 0; i < cells; i++) v[i*4 + j] = col[i];
        }
        if (a > 3) return zero();
        return detail::intern(op_t::spline_1d, x, nullptr, nullptr, 0.0, detail::intern_table(v),
                              spline_order {a, 0, 0}.pack(), shape->scale, shape->offset);
    }
///  c[i][j]: coefficient of u^i z^j, piecewise_2D nodes (or constants) on the same cells, folded in z.
    inline leaf_ptr spline_2d(const std::array<std::array<leaf_ptr, 4>, 4> &c, leaf_ptr r, leaf_ptr z,
                              const unsigned a=0, const unsigned b=0) {
        const leaf_node *shape = nullptr;
        for (auto &row : c) for (auto &n : row) if (n->op == op_t::piecewise_2d) shape = n.get();
        assert(shape && "spline_2d needs at least one table.");
        const size_t cells = shape->table->values.size();
        std::vector<double> v(cells*16);
        for (size_t i = 0; i < 4; i++) {
            for (size_t j = 0; j < 4; j++) {
                const auto col = detail::cell_values(c[i][j], cells);
                for (size_t k = 0; k < cells; k++) v[k*16 + i*4 + j] = col[k];
            }
        }
        if (a > 3 || b > 3) return zero();
        return detail::intern(op_t::spline_2d, r, z, nullptr, 0.0, detail::intern_table(v),
                              spline_order {a, b, shape->num_cols}.pack(), shape->scale, shape->offset);
    }
///  Another member of the family of `n`.
    inline leaf_ptr spline_member(const leaf_node *n, const unsigned a, const unsigned b) {
        if (a > 3 || b > 3) return zero();
        const spline_order o = spline_order::unpack(n->num_cols);
        return detail::intern(n->op, n->args[0], n->args[1], nullptr, 0.0, n->table, spline_order {a, b, o.columns}.pack(),
                              n->scale, n->offset);
    }
///  Host values (also the oracle of the device code in tests).
    inline double spline_1d_value(const leaf_node *n, const double x) {
        const spline_order o = spline_order::unpack(n->num_cols);
        const size_t cell = table_index(x, n->scale[0], n->offset[0], n->table->values.size()/4);
        return detail::cubic(&n->table->values[cell*4], x, o.a);
    }
    inline double spline_2d_value(const leaf_node *n, const double r, const double z) {
        const spline_order o = spline_order::unpack(n->num_cols);
        const size_t rows = n->table->values.size()/16/o.columns;
        const size_t cell = table_index(r, n->scale[0], n->offset[0], rows)*o.columns +
                            table_index(z, n->scale[1], n->offset[1], o.columns);
        const double *k = &n->table->values[cell*16];
        double row[4];
        for (size_t i = 0; i < 4; i++) row[i] = detail::cubic(k + 4*i, z, o.b);
        const double inv = 1.0/n->scale[0];
        const double u = (r - n->offset[0])*inv;
        double value = detail::cubic(row, u, o.a);
        for (unsigned i = 0; i < o.a; i++) value *= inv;
        return value;
    }

//------------------------------------------------------------------------------
//  Operators (double operands become constants).
//------------------------------------------------------------------------------
    inline leaf_ptr operator+(leaf_ptr l, leaf_ptr r) { return add(l, r); }
    inline leaf_ptr operator-(leaf_ptr l, leaf_ptr r) { return sub(l, r); }
    inline leaf_ptr operator*(leaf_ptr l, leaf_ptr r) { return mul(l, r); }
    inline leaf_ptr operator/(leaf_ptr l, leaf_ptr r) { return div(l, r); }
    inline leaf_ptr operator-(leaf_ptr a) { return mul(none(), a); }
    inline leaf_ptr operator+(const double l, leaf_ptr r) { return add(constant(l), r); }
    inline leaf_ptr operator+(leaf_ptr l, const double r) { return add(l, constant(r)); }
    inline leaf_ptr operator-(const double l, leaf_ptr r) { return sub(constant(l), r); }
    inline leaf_ptr operator-(leaf_ptr l, const double r) { return sub(l, constant(r)); }
    inline leaf_ptr operator*(const double l, leaf_ptr r) { return mul(constant(l), r); }
    inline leaf_ptr operator*(leaf_ptr l, const double r) { return mul(l, constant(r)); }
    inline leaf_ptr operator/(const double l, leaf_ptr r) { return div(constant(l), r); }
    inline leaf_ptr operator/(leaf_ptr l, const double r) { return div(l, constant(r)); }
    inline leaf_ptr fma(const double a, leaf_ptr b, leaf_ptr c) { return fma(constant(a), b, c); }
    inline leaf_ptr fma(leaf_ptr a, const double b, leaf_ptr c) { return fma(a, constant(b), c); }
    inline leaf_ptr fma(leaf_ptr a, leaf_ptr b, const double c) { return fma(a, b, constant(c)); }
    inline leaf_ptr pow(leaf_ptr a, const double b) { return pow(a, constant(b)); }
    inline leaf_ptr pow(const double a, leaf_ptr b) { return pow(constant(a), b); }
    inline leaf_ptr atan(const double x, leaf_ptr y) { return atan(constant(x), y); }
    inline leaf_ptr atan(leaf_ptr x, const double y) { return atan(x, constant(y)); }

//------------------------------------------------------------------------------
//  Table index (the contract of piecewise.hpp:26-65).
//------------------------------------------------------------------------------
    inline size_t table_index(const double x, const double scale, const double offset, const size_t n) {
        const double u = std::fmin(std::fmax((x - offset)/scale, 0.0), static_cast<double> (n - 1));
        return static_cast<size_t> (u);
    }

//------------------------------------------------------------------------------
//  evaluate().
//------------------------------------------------------------------------------
    namespace detail {
        inline const std::vector<double> &eval(leaf_node *n,
                                               std::unordered_map<const leaf_node *, std::vector<double>> &memo) {
            auto it = memo.find(n);
            if (it != memo.end()) return it->second;
            std::vector<double> out;
            auto bin = [&] (auto f) {
                const auto &a = eval(n->args[0].get(), memo);
                const auto &b = eval(n->args[1].get(), memo);
                const size_t s = std::max(a.size(), b.size());
                out.resize(s);
                for (size_t i = 0; i < s; i++) out[i] = f(a[a.size() == 1 ? 0 : i], b[b.size() == 1 ? 0 : i]);
            };
            auto un = [&] (auto f) {
                const auto &a = eval(n->args[0].get(), memo);
                out.resize(a.size());
                for (size_t i = 0; i < a.size(); i++) out[i] = f(a[i]);
            };
            switch (n->op) {
                case op_t::constant: out.assign(1, n->value); break;
                case op_t::variable: out = n->buffer; break;
                case op_t::pseudo: out = eval(n->args[0].get(), memo); break;
                case op_t::add: bin([] (double a, double b) { return a + b; }); break;
                case op_t::sub: bin([] (double a, double b) { return a - b; }); break;
                case op_t::mul: bin([] (double a, double b) { return a*b; }); break;
                case op_t::div: bin([] (double a, double b) { return a/b; }); break;
                case op_t::pow: bin([] (double a, double b) { return std::pow(a, b); }); break;
                case op_t::atan: bin([] (double x, double y) { return std::atan2(y, x); }); break;
                case op_t::sqrt: un([] (double a) { return std::sqrt(a); }); break;
                case op_t::exp: un([] (double a) { return std::exp(a); }); break;
                case op_t::log: un([] (double a) { return std::log(a); }); break;
                case op_t::erfi: un([] (double a) { return gfb::erfi(a); }); break;
                case op_t::nan_to_zero: un([] (double a) { return a == a ? a : 0.0; }); break;
                case op_t::sin: un([] (double a) { return std::sin(a); }); break;
                case op_t::cos: un([] (double a) { return std::cos(a); }); break;
                case op_t::fma: {
                    const auto &a = eval(n->args[0].get(), memo);
                    const auto &b = eval(n->args[1].get(), memo);
                    const auto &c = eval(n->args[2].get(), memo);
                    const size_t s = std::max(a.size(), std::max(b.size(), c.size()));
                    out.resize(s);
                    for (size_t i = 0; i < s; i++)
                        out[i] = std::fma(a[a.size() == 1 ? 0 : i], b[b.size() == 1 ? 0 : i], c[c.size() == 1 ? 0 : i]);
                    break;
                }
                case op_t::piecewise_1d: {
                    const auto &t = n->table->values;
                    un([&] (double a) { return t[table_index(a, n->scale[0], n->offset[0], t.size())]; });
                    break;
                }
                case op_t::index_1d: {
                    const auto &t = n->args[0]->buffer;
                    const auto &a = eval(n->args[1].get(), memo);
                    out.resize(a.size());
                    for (size_t i = 0; i < a.size(); i++) out[i] = t[table_index(a[i], n->scale[0], n->offset[0], t.size())];
                    break;
                }
                case op_t::index_2d: {
                    const auto &t = n->args[0]->buffer;
                    const auto &a = eval(n->args[1].get(), memo);
                    const auto &b = eval(n->args[2].get(), memo);
                    const size_t rows = t.size()/n->num_cols, sz = std::max(a.size(), b.size());
                    out.resize(sz);
                    for (size_t i = 0; i < sz; i++)
                        out[i] = t[table_index(a[a.size() == 1 ? 0 : i], n->scale[0], n->offset[0], rows)*n->num_cols +
                                   table_index(b[b.size() == 1 ? 0 : i], n->scale[1], n->offset[1], n->num_cols)];
                    break;
                }
                case op_t::fourier: {
                    const auto &a = eval(n->args[0].get(), memo);
                    const auto &b = eval(n->args[1].get(), memo);
                    const auto &c = eval(n->args[2].get(), memo);
                    const size_t sz = std::max(a.size(), std::max(b.size(), c.size()));
                    out.resize(sz);
                    const fourier_order o = fourier_order::unpack(n->num_cols);
                    for (size_t i = 0; i < sz; i++)
                        out[i] = fourier_value(*n->table, o, n->scale[0], n->offset[0], a[a.size() == 1 ? 0 : i],
                                               b[b.size() == 1 ? 0 : i], c[c.size() == 1 ? 0 : i]);
                    break;
                }
                case op_t::spline_1d: un([&] (double a) { return spline_1d_value(n, a); }); break;
                case op_t::spline_2d: bin([&] (double a, double b) { return spline_2d_value(n, a, b); }); break;
                case op_t::piecewise_2d: {
                    const auto &t = n->table->values;
                    const size_t rows = t.size()/n->num_cols;
                    bin([&] (double a, double b) {
                        return t[table_index(a, n->scale[0], n->offset[0], rows)*n->num_cols +
                                 table_index(b, n->scale[1], n->offset[1], n->num_cols)];
                    });
                    break;
                }
            }
            return memo.emplace(n, std::move(out)).first->second;
        }
    }
    inline backend::buffer<double> leaf_node::evaluate() {
        std::unordered_map<const leaf_node *, std::vector<double>> memo;
        return backend::buffer<double> (detail::eval(this, memo));
    }

//------------------------------------------------------------------------------
//  df(): forward symbolic differentiation, memoised per (node, x).
//------------------------------------------------------------------------------
    inline leaf_ptr leaf_node::df(leaf_ptr x) {
        if (x.get() == this) return one();
        if (op == op_t::constant || op == op_t::variable || op == op_t::pseudo || is_piecewise() || is_index()) return zero();
        auto &memo = detail::caches().df;
        const auto k = std::make_pair(static_cast<const leaf_node *> (this), static_cast<const leaf_node *> (x.get()));
        auto it = memo.find(k);
        if (it != memo.end()) return it->second;
        auto self = shared_from_this();
        const auto &a = args[0];
        const auto &b = args[1];
        leaf_ptr r;
        switch (op) {
            case op_t::add: r = add(a->df(x), b->df(x)); break;
            case op_t::sub: r = sub(a->df(x), b->df(x)); break;
            case op_t::mul: r = add(mul(a->df(x), b), mul(a, b->df(x))); break;
            case op_t::div: {
//  (a/b)' = a'/b - (a/b)*(b'/b); both shapes reuse the quotient node itself.
                auto da = a->df(x), db = b->df(x);
                if (db->is_constant(0.0)) r = div(da, b);
                else r = sub(div(da, b), mul(self, div(db, b)));
                break;
            }
            case op_t::fma: r = add(add(mul(a->df(x), b), mul(a, b->df(x))), args[2]->df(x)); break;
            case op_t::sqrt: r = div(a->df(x), mul(constant(2.0), self)); break;
            case op_t::exp: r = mul(self, a->df(x)); break;
            case op_t::log: r = div(a->df(x), a); break;
            case op_t::erfi: r = mul(mul(constant(2.0/std::sqrt(M_PI)), exp(mul(a, a))), a->df(x)); break;
            case op_t::nan_to_zero: r = nan_to_zero(a->df(x)); break;
            case op_t::pow: {
                if (b->is_constant()) {
                    r = mul(mul(b, pow(a, constant(b->value - 1.0))), a->df(x));
                } else {
                    r = mul(self, add(mul(b->df(x), log(a)), mul(b, div(a->df(x), a))));
                }
                break;
            }
            case op_t::sin: r = mul(cos(a), a->df(x)); break;
            case op_t::cos: r = mul(mul(none(), sin(a)), a->df(x)); break;
            case op_t::atan: {
//  d atan2(y, x) = (x dy - y dx)/(x^2 + y^2), args = (x, y).
                r = div(sub(mul(a, b->df(x)), mul(b, a->df(x))), add(mul(a, a), mul(b, b)));
                break;
            }
            case op_t::spline_1d: {
                const spline_order o = spline_order::unpack(num_cols);
                r = mul(spline_member(this, o.a + 1, 0), a->df(x));
                break;
            }
            case op_t::spline_2d: {
                const spline_order o = spline_order::unpack(num_cols);
                r = add(mul(spline_member(this, o.a + 1, o.b), a->df(x)), mul(spline_member(this, o.a, o.b + 1), b->df(x)));
                break;
            }
            case op_t::fourier: {
                const fourier_order o = fourier_order::unpack(num_cols);
                r = add(add(mul(fourier_series(table, a, b, args[2], scale[0], offset[0], {o.a + 1, o.b, o.c, o.base}), a->df(x)),
                            mul(fourier_series(table, a, b, args[2], scale[0], offset[0], {o.a, o.b + 1, o.c, o.base}), b->df(x))),
                        mul(fourier_series(table, a, b, args[2], scale[0], offset[0], {o.a, o.b, o.c + 1, o.base}), args[2]->df(x)));
                break;
            }
            default: r = zero(); break;
        }
        memo.emplace(k, r);
        return r;
    }

//------------------------------------------------------------------------------
//  Generic rebuild of a node with new arguments (used by remove_pseudo and
//  by substitution in the solvers).
//------------------------------------------------------------------------------
    inline leaf_ptr rebuild(const leaf_node *n, leaf_ptr a, leaf_ptr b, leaf_ptr c) {
        switch (n->op) {
            case op_t::add: return add(a, b);
            case op_t::sub: return sub(a, b);
            case op_t::mul: return mul(a, b);
            case op_t::div: return div(a, b);
            case op_t::fma: return fma(a, b, c);
            case op_t::sqrt: return sqrt(a);
            case op_t::exp: return exp(a);
            case op_t::log: return log(a);
            case op_t::erfi: return erfi(a);
            case op_t::nan_to_zero: return nan_to_zero(a);
            case op_t::pow: return pow(a, b);
            case op_t::sin: return sin(a);
            case op_t::cos: return cos(a);
            case op_t::atan: return atan(a, b);
            case op_t::pseudo: return pseudo_variable(a);
            case op_t::piecewise_1d: case op_t::piecewise_2d:
                return detail::intern(n->op, a, b, nullptr, 0.0, n->table, n->num_cols, n->scale, n->offset);
            case op_t::fourier: case op_t::index_1d: case op_t::index_2d: case op_t::spline_1d: case op_t::spline_2d:
                return detail::intern(n->op, a, b, c, 0.0, n->table, n->num_cols, n->scale, n->offset);
            default: return std::const_pointer_cast<leaf_node> (n->shared_from_this());
        }
    }

    inline leaf_ptr leaf_node::remove_pseudo() {
        if (op == op_t::constant || op == op_t::variable) return shared_from_this();
        auto &memo = detail::caches().no_pseudo;
        auto it = memo.find(this);
        if (it != memo.end()) return it->second;
        leaf_ptr r;
        if (op == op_t::pseudo) {
            r = args[0]->remove_pseudo();
        } else {
            leaf_ptr a = args[0].get() ? args[0]->remove_pseudo() : leaf_ptr();
            leaf_ptr b = args[1].get() ? args[1]->remove_pseudo() : leaf_ptr();
            leaf_ptr c = args[2].get() ? args[2]->remove_pseudo() : leaf_ptr();
            r = (a == args[0] && b == args[1] && c == args[2]) ? shared_from_this() : rebuild(this, a, b, c);
        }
        memo.emplace(this, r);
        return r;
    }

///  Substitute nodes (old -> new) throughout an expression.
    inline leaf_ptr substitute(leaf_ptr n, const std::unordered_map<const leaf_node *, leaf_ptr> &map,
                               std::unordered_map<const leaf_node *, leaf_ptr> &memo) {
        auto f = map.find(n.get());
        if (f != map.end()) return f->second;
        if (n->op == op_t::constant || n->op == op_t::variable) return n;
        auto it = memo.find(n.get());
        if (it != memo.end()) return it->second;
        leaf_ptr a = n->args[0].get() ? substitute(n->args[0], map, memo) : leaf_ptr();
        leaf_ptr b = n->args[1].get() ? substitute(n->args[1], map, memo) : leaf_ptr();
        leaf_ptr c = n->args[2].get() ? substitute(n->args[2], map, memo) : leaf_ptr();
        leaf_ptr r = (a == n->args[0] && b == n->args[1] && c == n->args[2]) ? n : rebuild(n.get(), a, b, c);
        memo.emplace(n.get(), r);
        return r;
    }

//------------------------------------------------------------------------------
///  Reverse-mode symbolic gradient: d f / d x_i for all requested leaves in ONE backward sweep
///  over the DAG.  Mathematically identical to {f->df(x_i)}; the expression shape differs:
///  every interior adjoint  df/dn  is formed once and shared by all x_i, where forward mode
///  carries one tangent per x_i through every node.  For the ray Hamiltonian (7 derivatives of
///  one scalar whose coordinate dependence funnels through psi(R, Z)) this is the cheaper form.
///  Semantics match df(): coefficients of piecewise nodes and pseudo variables are leaves.
//------------------------------------------------------------------------------
    inline std::vector<leaf_ptr> gradient(leaf_ptr f, const std::vector<leaf_ptr> &xs) {
        std::vector<leaf_node *> order;
        std::unordered_map<const leaf_node *, bool> seen;
        std::function<void(leaf_node *)> visit = [&] (leaf_node *n) {
            if (seen[n]) return;
            seen[n] = true;
            if (n->op != op_t::pseudo && !n->is_piecewise() && !n->is_index()) {
                for (size_t i = 0, ie = n->num_args(); i < ie; i++) visit(n->args[i].get());
            }
            order.push_back(n);
        };
        visit(f.get());
        std::unordered_map<const leaf_node *, leaf_ptr> adj;
        auto accumulate = [&adj] (const leaf_ptr &n, leaf_ptr v) {
            if (n->is_constant()) return;
            auto it = adj.find(n.get());
            if (it == adj.end()) adj.emplace(n.get(), v);
            else it->second = add(it->second, v);
        };
        adj.emplace(f.get(), one());
        for (auto it = order.rbegin(); it != order.rend(); ++it) {
            leaf_node *n = *it;
            auto found = adj.find(n);
            if (found == adj.end()) continue;
            const leaf_ptr a = found->second;
            if (a->is_constant(0.0)) continue;
            const leaf_ptr self = n->shared_from_this();
            const leaf_ptr &x = n->args[0];
            const leaf_ptr &y = n->args[1];
            switch (n->op) {
                case op_t::add: accumulate(x, a); accumulate(y, a); break;
                case op_t::sub: accumulate(x, a); accumulate(y, mul(none(), a)); break;
                case op_t::mul: accumulate(x, mul(a, y)); accumulate(y, mul(a, x)); break;
                case op_t::div: {
                    auto q = div(a, y);
                    accumulate(x, q);
                    accumulate(y, mul(none(), mul(q, self)));
                    break;
                }
                case op_t::fma: accumulate(x, mul(a, y)); accumulate(y, mul(a, x)); accumulate(n->args[2], a); break;
                case op_t::sqrt: accumulate(x, div(a, mul(constant(2.0), self))); break;
                case op_t::exp: accumulate(x, mul(a, self)); break;
                case op_t::log: accumulate(x, div(a, x)); break;
                case op_t::erfi: accumulate(x, mul(a, mul(constant(2.0/std::sqrt(M_PI)), exp(mul(x, x))))); break;
                case op_t::nan_to_zero: accumulate(x, a); break;
                case op_t::pow:
                    if (y->is_constant()) {
                        accumulate(x, mul(a, mul(y, pow(x, constant(y->value - 1.0)))));
                    } else {
                        accumulate(x, mul(a, mul(y, div(self, x))));
                        accumulate(y, mul(a, mul(self, log(x))));
                    }
                    break;
                case op_t::sin: accumulate(x, mul(a, cos(x))); break;
                case op_t::cos: accumulate(x, mul(none(), mul(a, sin(x)))); break;
                case op_t::atan: {
                    auto q = div(a, add(mul(x, x), mul(y, y)));
                    accumulate(x, mul(none(), mul(q, y)));
                    accumulate(y, mul(q, x));
                    break;
                }
                case op_t::spline_1d: {
                    const spline_order o = spline_order::unpack(n->num_cols);
                    accumulate(x, mul(a, spline_member(n, o.a + 1, 0)));
                    break;
                }
                case op_t::spline_2d: {
                    const spline_order o = spline_order::unpack(n->num_cols);
                    accumulate(x, mul(a, spline_member(n, o.a + 1, o.b)));
                    accumulate(y, mul(a, spline_member(n, o.a, o.b + 1)));
                    break;
                }
                case op_t::fourier: {
                    const fourier_order o = fourier_order::unpack(n->num_cols);
                    const leaf_ptr &v = n->args[2];
                    accumulate(x, mul(a, fourier_series(n->table, x, y, v, n->scale[0], n->offset[0], {o.a + 1, o.b, o.c, o.base})));
                    accumulate(y, mul(a, fourier_series(n->table, x, y, v, n->scale[0], n->offset[0], {o.a, o.b + 1, o.c, o.base})));
                    accumulate(v, mul(a, fourier_series(n->table, x, y, v, n->scale[0], n->offset[0], {o.a, o.b, o.c + 1, o.base})));
                    break;
                }
                default: break;
            }
        }
        std::vector<leaf_ptr> result;
        for (auto &x : xs) {
            auto it = adj.find(x.get());
            result.push_back(it == adj.end() ? zero() : it->second);
        }
        return result;
    }

    inline bool leaf_node::is_constant_like() {
        if (op == op_t::constant) return true;
        if (op == op_t::variable) return false;
        for (size_t i = 0, ie = num_args(); i < ie; i++) if (!args[i]->is_constant_like()) return false;
        return true;
    }

    inline std::string leaf_node::to_string() {
        static const char *names[] = {"const", "var", "pseudo", "+", "-", "*", "/", "fma", "sqrt", "exp", "log",
                                      "pow", "sin", "cos", "atan", "pw1d", "pw2d", "fourier", "erfi", "nan0", "idx1d", "idx2d",
                                      "spline1d", "spline2d"};
        std::ostringstream s;
        s.precision(17);
        if (op == op_t::constant) { s << value; return s.str(); }
        if (op == op_t::variable) return symbol;
        s << "(" << names[static_cast<int> (op)];
        for (size_t i = 0, ie = num_args(); i < ie; i++) s << " " << args[i]->to_string();
        s << ")";
        return s.str();
    }
}

#endif /* gfb_graph_node_hpp */
