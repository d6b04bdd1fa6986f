//------------------------------------------------------------------------------
//  emit.hpp -- CUDA code generation for the sm_100a skeletons.
//
//  Turns the interned expression DAG of one work item into a traits struct for
//  skeleton.cuh: a straight-line `body` (one SSA statement per live node),
//  load/apply/store glue and the packed spline-table groups it reads.
//
//  Reference counterpart: the per-node compile() methods plus
//  cuda_context::create_kernel_prefix/postfix
//  (/root/reference/graph_framework/cuda_context.hpp:713-946,
//   arithmetic.hpp:645-669 etc., piecewise.hpp:348-437, 1071-1208).
//  Differences that matter for speed:
//    * a reciprocal is computed once per distinct denominator and multiplied;
//    * the table index is computed once per distinct (argument, grid);
//    * all coefficient tables that share an index are packed cell-major
//      (array-of-structs) so one cell is a contiguous row, read with 16-byte
//      loads from L1/L2 or staged into shared memory by TMA when small;
//    * the body is a device function, not a whole kernel: the Runge-Kutta
//      stage loop calls ONE right-hand-side body instead of four inlined copies.
//------------------------------------------------------------------------------
#ifndef gfb_graph_emit_hpp
#define gfb_graph_emit_hpp

#include <algorithm>
#include <cstdio>
#include <set>
#include <unordered_set>

#include "node.hpp"

namespace jit {
    enum class kernel_kind { generic, rk2, rk4, newton };

//------------------------------------------------------------------------------
///  One packed group of coefficient tables sharing a cell index.
//------------------------------------------------------------------------------
    struct table_group {
        graph::op_t op;
        graph::leaf_ptr arg0, arg1;
        size_t num_cols;
        std::array<double, 2> scale, offset;
        size_t cells;
        std::vector<graph::table_ptr> members;
        size_t stride;                  ///< doubles per cell row (padded)
        bool staged;                    ///< copied to shared memory by TMA
        size_t smem_offset;             ///< byte offset in dynamic shared memory
        std::vector<double> packed;     ///< cells*stride doubles, cell major
        bool raw = false;               ///< packed was filled by the emitter (Fourier tables), do not repack
        int alias_input = -1;           ///< >= 0: no table of its own, the pointer of that kernel input (index_1D/2D)
        bool mode_table = false;        ///< the small [mode][xm, xn, row start, 0] table of a Fourier loop
///  Cubic-spline families (graph::spline_1d / spline_2d): one group per (table, arguments); `orders` are
///  the derivative orders (a, b) this kernel needs; all of them come out of ONE pass over the cell's row.
        bool spline = false;
        graph::table_ptr spline_table;
        std::set<std::pair<unsigned, unsigned>> orders;
        std::vector<const graph::leaf_node *> spline_nodes;     ///< in scan order (deterministic kernel text)
        bool emitted = false;
        size_t bytes() const { return packed.size()*sizeof(double); }
    };

//------------------------------------------------------------------------------
///  One device loop over Fourier modes: every graph::fourier_series node that shares the angles
///  (u, v) and the mode numbers is a member; `sets` are the coefficient tables involved.
//------------------------------------------------------------------------------
    struct fourier_set {
        graph::table_ptr table;
        const graph::leaf_node *s;
        double scale, offset;
        size_t group;                                       ///< pointer slot of the coefficient buffer
        std::vector<const graph::leaf_node *> members;
    };
    struct fourier_loop {
        const graph::leaf_node *u, *v;
        std::vector<double> xm, xn;
        size_t mode_group;                                  ///< pointer slot of the [mode][4] table
        double xn_step = 0.0;                               ///< xn increment inside a row of equal xm (0: no rows)
        std::vector<bool> row_start;
        std::vector<std::pair<unsigned, unsigned>> factors; ///< (b, c) with b + c >= 2: xm^b (-xn)^c tabulated per mode
        size_t stride = 4;                                  ///< doubles per mode: xm, -xn, row start, factors...
        std::vector<fourier_set> sets;
        bool emitted = false;
    };

//------------------------------------------------------------------------------
///  Everything the runtime needs to launch one emitted kernel.
//------------------------------------------------------------------------------
    struct kernel_info {
        std::string name;
        kernel_kind kind;
        graph::input_nodes<> inputs;
        graph::output_nodes<> outputs;          ///< one device buffer each, after the inputs
        std::vector<bool> input_written;
        std::vector<bool> input_loaded;         ///< false: only reached through index_1D/2D, never loaded per ray
        bool indexed_written = false;           ///< an indexed input is also a setter target: steps must not be fused
        std::vector<table_group> groups;        ///< pointer slots after the outputs (alias groups own none)
        size_t size;
        size_t smem_bytes;
        size_t num_statements;
        size_t num_divides;
        size_t num_reciprocals;
        bool has_mode_loop = false;             ///< the body contains a device loop over Fourier modes
    };

    struct emit_options {
        size_t stage_limit_bytes = 32*1024;     ///< largest group staged into shared memory
        size_t stage_total_bytes = 96*1024;
        bool share_reciprocals = true;
        bool stage_tables = true;
///  Refined hardware reciprocal / rsqrt seeds instead of IEEE division and sqrt, and a
///  multiplication by 1/scale in table indices (what -ffast-math does to the reference's kernels).
        bool fast_division = true;
        unsigned mode_loop_unroll = 2;          ///< unroll factor of Fourier mode loops (graph::fourier_series)
///  Fourier loops: inside a row of modes with equal m and equally spaced n, step the angle by one
///  rotation (4 FMAs) instead of a sincos per mode; each row starts from a fresh sincos.
        bool mode_recurrence = true;
///  Expressions that depend only on inputs the kernel never writes (the wave frequency of a ray: 1/w,
///  w^2, 1/w^2 ...) are evaluated once per launch, before the step loop, instead of in every stage.
        bool hoist_invariants = true;
        size_t unroll_stages_below = 640;       ///< unroll the RK stage loop for bodies up to this many statements
        unsigned block_size = 128;
///  Resident blocks per SM promised to ptxas; 0 = let the device layer pick the highest
///  value (4, 3, 2, 1) that compiles without register spills (-DGFB_MIN_BLOCKS).
        unsigned min_blocks = 0;
    };

    inline std::string literal(const double d) {
        char buf[64];
        if (std::isinf(d)) return d > 0 ? "(1.0/0.0)" : "(-1.0/0.0)";
        if (std::isnan(d)) return "(0.0/0.0)";
        std::snprintf(buf, sizeof(buf), "%.17g", d);
        std::string s(buf);
        if (s.find_first_of(".en") == std::string::npos) s += ".0";
        if (d < 0.0 || (d == 0.0 && std::signbit(d))) s = "(" + s + ")";
        return s;
    }

//------------------------------------------------------------------------------
///  Emits one kernel.
//------------------------------------------------------------------------------
    class body_emitter {
    private:
        std::ostringstream &out;
        const emit_options &opt;
        kernel_info &info;
        std::unordered_map<const graph::leaf_node *, std::string> reg;
        std::unordered_map<const graph::leaf_node *, std::string> inv_reg;
        std::unordered_map<const graph::leaf_node *, size_t> denominators;
        std::unordered_map<const graph::leaf_node *, std::pair<size_t, size_t>> slot;   // piecewise -> (group, member)
        std::unordered_set<const graph::leaf_node *> visited;
        std::vector<std::string> index_reg;
        std::vector<std::vector<bool>> pair_loaded;
        std::vector<fourier_loop> floops;
        std::unordered_map<const graph::leaf_node *, size_t> floop_of;
///  Arguments that occur under both sin and cos in this kernel: one sincos() serves both.
        std::unordered_map<const graph::leaf_node *, std::pair<const graph::leaf_node *, const graph::leaf_node *>> trig;
///  Launch invariants: their statements go to K::invariants(), the body reads h[k].  While they are being
///  emitted `out`, `reg` and `inv_reg` hold the invariant function's text and names.
        std::unordered_map<const graph::leaf_node *, bool> invariant_memo;
        std::unordered_map<const graph::leaf_node *, std::string> reg_other, inv_reg_other;
        std::string text_other;
        bool in_invariants = false;
        size_t num_invariants = 0;

        bool is_invariant(const graph::leaf_node *n) {
            n = strip(n);
            auto it = invariant_memo.find(n);
            if (it != invariant_memo.end()) return it->second;
            bool result = true;
            if (n->op == graph::op_t::variable) {
                result = false;
                for (size_t i = 0; i < info.inputs.size(); i++)
                    if (info.inputs[i].get() == n) result = !info.input_written[i];
            } else if (n->is_piecewise() || n->is_spline() || n->is_index() || n->op == graph::op_t::fourier) {
                result = false;         // table look-ups stay in the body: their row pointers are per context
            } else {
                for (size_t i = 0, ie = n->num_args(); i < ie && result; i++) result = is_invariant(n->args[i].get());
            }
            invariant_memo.emplace(n, result);
            return result;
        }
        void switch_context() {
            std::string mine = out.str();
            out.str(text_other);
            out.seekp(0, std::ios_base::end);
            text_other.swap(mine);
            reg.swap(reg_other);
            inv_reg.swap(inv_reg_other);
            in_invariants = !in_invariants;
        }
///  Body context: the name under which the body reads invariant node n (or its reciprocal).
        std::string hoisted(const graph::leaf_node *n, const bool reciprocal) {
            switch_context();
            std::string local = emit(n);
            if (reciprocal) {
                auto f = inv_reg.find(n);
                if (f == inv_reg.end()) {
                    const std::string iname = "i" + local_id(n);
                    out << "        const double " << iname << " = "
                        << (opt.fast_division ? "gfb::rcp(" + local + ")" : "1.0/" + local) << ";" << std::endl;
                    info.num_reciprocals++;
                    f = inv_reg.emplace(n, iname).first;
                }
                local = f->second;
            }
            const std::string slot_name = "h[" + std::to_string(num_invariants++) + "]";
            out << "        " << slot_name << " = " << local << ";" << std::endl;
            switch_context();
            return slot_name;
        }

///  Register names count nodes in the order this kernel first names them, not by the thread-wide creation
///  counter: the text of a kernel (and its sha256, which ties ncu captures to binaries) must not depend on
///  what else the thread built before.
        std::unordered_map<const graph::leaf_node *, size_t> local_ids;
        std::string local_id(const graph::leaf_node *n) {
            return std::to_string(local_ids.emplace(n, local_ids.size()).first->second);
        }

        static const graph::leaf_node *strip(const graph::leaf_node *n) {
            while (n->op == graph::op_t::pseudo) n = n->args[0].get();
            return n;
        }

//  Pass 1: find table groups and count how often each node is a denominator.
        void scan(const graph::leaf_node *n) {
            n = strip(n);
            if (!visited.insert(n).second) return;
            if (n->op == graph::op_t::variable) {
                bool found = false;
                for (auto &in : info.inputs) found = found || in.get() == n;
                if (!found) {
                    std::cerr << "Kernel " << info.name << ": variable " << n->symbol
                              << " is used but is not an input." << std::endl;
                    std::exit(-1);
                }
            }
            if (n->op == graph::op_t::div) denominators[strip(n->args[1].get())]++;
            if (n->op == graph::op_t::sin) trig[strip(n->args[0].get())].first = n;
            if (n->op == graph::op_t::cos) trig[strip(n->args[0].get())].second = n;
            if (n->op == graph::op_t::fourier) {
                const graph::leaf_node *fs = strip(n->args[0].get());
                const graph::leaf_node *fu = strip(n->args[1].get());
                const graph::leaf_node *fv = strip(n->args[2].get());
                size_t l = 0;
                for (; l < floops.size(); l++)
                    if (floops[l].u == fu && floops[l].v == fv && floops[l].xm == n->table->xm && floops[l].xn == n->table->xn) break;
                if (l == floops.size()) {
                    fourier_loop loop;
                    loop.u = fu;
                    loop.v = fv;
                    loop.xm = n->table->xm;
                    loop.xn = n->table->xn;
//  Rows: runs of equal xm whose xn advance by one common step (VMEC orders its modes that way).
                    std::vector<bool> row_start(loop.xm.size(), true);
                    if (opt.mode_recurrence) {
                        for (size_t m = 1; m < loop.xm.size() && loop.xn_step == 0.0; m++)
                            if (loop.xm[m] == loop.xm[m - 1] && loop.xn[m] != loop.xn[m - 1]) loop.xn_step = loop.xn[m] - loop.xn[m - 1];
                        size_t continued = 0;
                        for (size_t m = 1; m < loop.xm.size(); m++) {
                            row_start[m] = !(loop.xn_step != 0.0 && loop.xm[m] == loop.xm[m - 1] &&
                                             loop.xn[m] - loop.xn[m - 1] == loop.xn_step);
                            continued += row_start[m] ? 0 : 1;
                        }
                        if (3*continued < 2*loop.xm.size()) {       // too few modes would profit
                            loop.xn_step = 0.0;
                            row_start.assign(loop.xm.size(), true);
                        }
                    }
                    loop.row_start = row_start;
                    table_group g;                  // filled by finish_mode_tables() once every member is known
                    g.op = graph::op_t::fourier;
                    g.num_cols = 0;
                    g.cells = loop.xm.size();
                    g.stride = 4;
                    g.raw = true;
                    g.mode_table = true;
                    loop.mode_group = info.groups.size();
                    info.groups.push_back(g);
                    floops.push_back(loop);
                }
                auto &sets = floops[l].sets;
                size_t j = 0;
                for (; j < sets.size(); j++)
                    if (sets[j].table == n->table && sets[j].s == fs && sets[j].scale == n->scale[0] && sets[j].offset == n->offset[0]) break;
                if (j == sets.size()) {
                    fourier_set set;
                    set.table = n->table;
                    set.s = fs;
                    set.scale = n->scale[0];
                    set.offset = n->offset[0];
                    table_group g;
                    g.op = graph::op_t::fourier;
                    g.num_cols = 0;
                    g.cells = n->table->cells;
                    g.stride = 4;
                    g.raw = true;
                    g.packed.assign(n->table->values.begin(),
                                    n->table->values.begin() + n->table->xm.size()*n->table->cells*4);
                    set.group = info.groups.size();
                    info.groups.push_back(g);
                    sets.push_back(set);
                }
                sets[j].members.push_back(n);
                floop_of[n] = l;
            }
            if (n->is_index()) {
                const graph::leaf_node *var = n->args[0].get();
                int j = -1;
                for (size_t i = 0; i < info.inputs.size(); i++) if (info.inputs[i].get() == var) j = static_cast<int> (i);
                if (j < 0) {
                    std::cerr << "Kernel " << info.name << ": indexed variable " << var->symbol
                              << " is not an input." << std::endl;
                    std::exit(-1);
                }
                size_t g = 0;
                for (; g < info.groups.size(); g++) if (info.groups[g].alias_input == j) break;
                if (g == info.groups.size()) {
                    table_group grp;
                    grp.op = n->op;
                    grp.num_cols = 0;
                    grp.cells = var->size();
                    grp.stride = 1;
                    grp.staged = false;
                    grp.smem_offset = 0;
                    grp.raw = true;
                    grp.alias_input = j;
                    info.groups.push_back(grp);
                }
                slot[n] = {g, 0};
                if (info.input_written[j]) info.indexed_written = true;
//  The array itself is not a per-ray value: only the index arguments are expressions.
                scan(n->args[1].get());
                if (n->op == graph::op_t::index_2d) scan(n->args[2].get());
                return;
            }
            if (n->is_spline()) {
                const graph::leaf_node *a0 = strip(n->args[0].get());
                const graph::leaf_node *a1 = n->args[1].get() ? strip(n->args[1].get()) : nullptr;
                const graph::spline_order o = graph::spline_order::unpack(n->num_cols);
                size_t g = 0;
                for (; g < info.groups.size(); g++) {
                    auto &grp = info.groups[g];
                    if (grp.spline && grp.op == n->op && grp.spline_table == n->table && grp.arg0.get() == a0 &&
                        grp.arg1.get() == a1 && grp.scale == n->scale && grp.offset == n->offset) break;
                }
                if (g == info.groups.size()) {
                    table_group grp;
                    grp.op = n->op;
                    grp.arg0 = std::const_pointer_cast<graph::leaf_node> (a0->shared_from_this());
                    if (a1) grp.arg1 = std::const_pointer_cast<graph::leaf_node> (a1->shared_from_this());
                    grp.num_cols = o.columns;
                    grp.scale = n->scale;
                    grp.offset = n->offset;
                    grp.stride = n->op == graph::op_t::spline_2d ? 16 : 4;
                    grp.cells = n->table->values.size()/grp.stride;
                    grp.raw = true;
                    grp.spline = true;
                    grp.spline_table = n->table;
                    grp.packed = n->table->values;
                    info.groups.push_back(grp);
                }
                info.groups[g].orders.insert({o.a, o.b});
                info.groups[g].spline_nodes.push_back(n);
                slot[n] = {g, 0};
            }
            if (n->is_piecewise()) {
                const graph::leaf_node *a0 = strip(n->args[0].get());
                const graph::leaf_node *a1 = n->args[1].get() ? strip(n->args[1].get()) : nullptr;
                size_t g = 0;
                for (; g < info.groups.size(); g++) {
                    auto &grp = info.groups[g];
                    if (!grp.spline && grp.op == n->op && grp.arg0.get() == a0 && grp.arg1.get() == a1 &&
                        grp.num_cols == n->num_cols && grp.scale == n->scale && grp.offset == n->offset &&
                        grp.cells == n->table->values.size()) break;
                }
                if (g == info.groups.size()) {
                    table_group grp;
                    grp.op = n->op;
                    grp.arg0 = std::const_pointer_cast<graph::leaf_node> (a0->shared_from_this());
                    if (a1) grp.arg1 = std::const_pointer_cast<graph::leaf_node> (a1->shared_from_this());
                    grp.num_cols = n->num_cols;
                    grp.scale = n->scale;
                    grp.offset = n->offset;
                    grp.cells = n->table->values.size();
                    info.groups.push_back(grp);
                }
                auto &members = info.groups[g].members;
                size_t m = 0;
                for (; m < members.size(); m++) if (members[m] == n->table) break;
                if (m == members.size()) members.push_back(n->table);
                slot[n] = {g, m};
            }
            for (size_t i = 0, ie = n->num_args(); i < ie; i++) scan(n->args[i].get());
        }

//  Per mode: xm, -xn, row-start flag, then xm^b (-xn)^c for every derivative order (b, c) with
//  b + c >= 2 that a member of the loop needs -- so that a weight is ONE multiply with the trig value.
        void finish_mode_tables() {
            info.has_mode_loop = !floops.empty();
            for (auto &loop : floops) {
                std::set<std::pair<unsigned, unsigned>> need;
                for (auto &set : loop.sets) {
                    for (auto *m : set.members) {
                        const graph::fourier_order o = graph::fourier_order::unpack(m->num_cols);
                        if (o.b + o.c >= 2) need.insert({o.b, o.c});
                    }
                }
                loop.factors.assign(need.begin(), need.end());
                loop.stride = (3 + loop.factors.size() + 1)/2*2;
                table_group &g = info.groups[loop.mode_group];
                g.stride = loop.stride;
                g.packed.assign(loop.xm.size()*loop.stride, 0.0);
                for (size_t m = 0; m < loop.xm.size(); m++) {
                    double *row = &g.packed[m*loop.stride];
                    row[0] = loop.xm[m];
                    row[1] = -loop.xn[m];
                    row[2] = loop.row_start[m] ? 1.0 : 0.0;
                    for (size_t f = 0; f < loop.factors.size(); f++) {
                        double v = 1.0;
                        for (unsigned i = 0; i < loop.factors[f].first; i++) v *= loop.xm[m];
                        for (unsigned i = 0; i < loop.factors[f].second; i++) v *= -loop.xn[m];
                        row[3 + f] = v;
                    }
                }
            }
        }

        void pack_groups() {
            size_t staged_total = 0;
            size_t offset = 16;     // mbarrier lives in the first 16 bytes
            for (auto &g : info.groups) {
                if (g.alias_input >= 0) continue;
                if (g.raw) {
//  Fourier tables keep their layout; the small [mode][4] mode-number table is staged.
                    if (g.packed.size() & 1) g.packed.push_back(0.0);
                    g.staged = opt.stage_tables && (g.mode_table || g.spline) && g.bytes() <= opt.stage_limit_bytes &&
                               staged_total + g.bytes() <= opt.stage_total_bytes;
                    if (g.staged) {
                        g.smem_offset = offset;
                        offset += (g.bytes() + 127)/128*128;
                        staged_total += g.bytes();
                    } else {
                        g.smem_offset = 0;
                    }
                    continue;
                }
                const size_t k = g.members.size();
                const size_t bytes_even = g.cells*(k + (k & 1))*sizeof(double);
                g.staged = opt.stage_tables && bytes_even <= opt.stage_limit_bytes &&
                           staged_total + bytes_even <= opt.stage_total_bytes;
//  Shared memory rows get an odd stride: consecutive cells then start in
//  different banks, so 32 rays in 32 different cells read one member without
//  conflicts beyond the 2 wavefronts 64-bit accesses need anyway.  Global rows
//  get an even stride so every pair of members is one aligned 16-byte load.
                g.stride = g.staged ? (k | 1) : (k + (k & 1));
                g.packed.assign(((g.cells*g.stride + 1)/2)*2, 0.0);
                for (size_t m = 0; m < k; m++)
                    for (size_t c = 0; c < g.cells; c++)
                        g.packed[c*g.stride + m] = g.members[m]->values[c];
                if (g.staged) {
                    g.smem_offset = offset;
                    offset += (g.bytes() + 127)/128*128;
                    staged_total += g.bytes();
                } else {
                    g.smem_offset = 0;
                }
            }
            info.smem_bytes = staged_total ? offset : 0;
        }

        std::string index_expr(const std::string &x, const double scale, const double offset, const size_t n) {
//  The reference's contract: (uint)min(max((x - offset)/scale, 0), n - 1)  (piecewise.hpp:26-65).  The clamp
//  is done on the integer side: the saturating conversion cvt.rzi.u32.f64 already maps everything below
//  zero (and NaN, which the reference's max(x, 0) also turns into 0) to 0, so min(trunc_sat(u), n - 1) is the
//  same cell for every double u -- one FP64-pipe instruction instead of two compares and a conversion.
            const std::string u = opt.fast_division ? "(" + x + " - " + literal(offset) + ")*" + literal(1.0/scale)
                                                    : "(" + x + " - " + literal(offset) + ")/" + literal(scale);
            return "min(__double2uint_rz(" + u + "), " + std::to_string(n - 1) + "u)";
        }

        std::string group_pointer(const size_t g) {
            auto &grp = info.groups[g];
            if (grp.staged) return "reinterpret_cast<const double *> (gfb::smem + " + std::to_string(grp.smem_offset) + ")";
            return "tg[" + std::to_string(g) + "]";
        }

//  One loop over the modes computes every member of the loop group (see graph::fourier_series).
        void emit_fourier_loop(fourier_loop &loop) {
            if (loop.emitted) return;
            loop.emitted = true;
            using graph::fourier_order;
            const std::string ureg = emit(loop.u), vreg = emit(loop.v);
            std::vector<std::string> sreg;
            for (auto &set : loop.sets) sreg.push_back(emit(set.s));
            const size_t id = loop.mode_group;
            for (auto &set : loop.sets) {
                for (auto *m : set.members) {
                    if (reg.count(m)) continue;
                    const std::string name = "t" + local_id(m);
                    out << "        double " << name << " = 0.0;" << std::endl;
                    reg.emplace(m, name);
                }
            }
            out << "        {" << std::endl
                << "            const double *mn" << id << " = " << group_pointer(loop.mode_group) << ";" << std::endl;
            for (size_t j = 0; j < loop.sets.size(); j++) {
                auto &set = loop.sets[j];
                out << "            const double *fc" << id << "_" << j << " = tg[" << set.group << "] + ("
                    << index_expr(sreg[j], set.scale, set.offset, set.table->cells) << ")*4u;" << std::endl;
            }
//  Which polynomial orders, and which (b, c, trig phase) weights are needed.
            std::set<std::pair<size_t, unsigned>> polys;
            std::set<std::array<unsigned, 3>> weights;
            for (size_t j = 0; j < loop.sets.size(); j++) {
                for (auto *m : loop.sets[j].members) {
                    const fourier_order o = fourier_order::unpack(m->num_cols);
                    polys.insert({j, o.a});
                    weights.insert({o.b, o.c, (o.b + o.c + (o.base ? 3u : 0u)) & 3u});
                }
            }
            const bool rows = loop.xn_step != 0.0;
            const std::string st = std::to_string(loop.stride);
            if (rows) {
//  Inside a row the angle m u - n v drops by xn_step*v per mode: one rotation instead of a sincos.
//  The row-start test is uniform across the warp (every lane is at the same mode).
                out << "            double rs" << id << ", rc" << id << ", sn = 0.0, cs = 1.0;" << std::endl
                    << "            sincos(" << literal(loop.xn_step) << "*" << vreg << ", &rs" << id << ", &rc" << id << ");" << std::endl;
            }
//  Cubic and its derivatives with the powers of s hoisted out of the mode loop:
//  P' = c1 + c2 (2 s) + c3 (3 s^2),  P''/2 = c2 + c3 (3 s),  P'''/6 = c3; the factors 2 and 6 are
//  applied to the accumulated sums after the loop.
            for (size_t j = 0; j < loop.sets.size(); j++) {
                bool d1 = false, d2 = false;
                for (auto &[jj, a] : polys) if (jj == j) { d1 = d1 || a == 1; d2 = d2 || a == 2; }
                const std::string &x = sreg[j];
                if (d1) out << "            const double x2_" << id << "_" << j << " = " << x << " + " << x << ", x3q_" << id << "_" << j
                            << " = 3.0*" << x << "*" << x << ";" << std::endl;
                if (d2) out << "            const double x3_" << id << "_" << j << " = 3.0*" << x << ";" << std::endl;
            }
            out << "#pragma unroll " << opt.mode_loop_unroll << std::endl
                << "            for (int m = 0; m < " << loop.xm.size() << "; m++) {" << std::endl
                << "                const double xm = mn" << id << "[" << st << "*m], nxn = mn" << id << "[" << st << "*m + 1];" << std::endl;
            if (rows) {
                out << "                if (mn" << id << "[" << st << "*m + 2] != 0.0) {" << std::endl
                    << "                    sincos(fma(nxn, " << vreg << ", xm*" << ureg << "), &sn, &cs);" << std::endl
                    << "                } else {" << std::endl
                    << "                    const double turned = fma(cs, rc" << id << ", sn*rs" << id << ");" << std::endl
                    << "                    sn = fma(sn, rc" << id << ", -(cs*rs" << id << "));" << std::endl
                    << "                    cs = turned;" << std::endl
                    << "                }" << std::endl;
            } else {
                out << "                double sn, cs;" << std::endl
                    << "                sincos(fma(nxn, " << vreg << ", xm*" << ureg << "), &sn, &cs);" << std::endl;
            }
            for (size_t j = 0; j < loop.sets.size(); j++) {
                const std::string c = "c" + std::to_string(j);
                const std::string &x = sreg[j];
                const std::string sj = std::to_string(id) + "_" + std::to_string(j);
                out << "                const double2 " << c << "lo = __ldg(reinterpret_cast<const double2 *> (fc" << id << "_" << j
                    << " + m*" << loop.sets[j].table->cells*4 << ")), " << c << "hi = __ldg(reinterpret_cast<const double2 *> (fc"
                    << id << "_" << j << " + m*" << loop.sets[j].table->cells*4 << ") + 1);" << std::endl;
                for (auto &[jj, a] : polys) {
                    if (jj != j) continue;
                    out << "                const double p" << j << "_" << a << " = ";
                    switch (a) {
                        case 0: out << "fma(fma(fma(" << c << "hi.y, " << x << ", " << c << "hi.x), " << x << ", " << c << "lo.y), " << x << ", " << c << "lo.x)"; break;
                        case 1: out << "fma(" << c << "hi.y, x3q_" << sj << ", fma(" << c << "hi.x, x2_" << sj << ", " << c << "lo.y))"; break;
                        case 2: out << "fma(" << c << "hi.y, x3_" << sj << ", " << c << "hi.x)"; break;
                        default: out << c << "hi.y"; break;
                    }
                    out << ";" << std::endl;
                }
            }
            for (auto &w : weights) {
                out << "                const double w" << w[0] << "_" << w[1] << "_" << w[2] << " = ";
                static const char *trig[4] = {"cs", "(-sn)", "(-cs)", "sn"};
                if (w[0] + w[1] == 0) {
                    out << trig[w[2]];
                } else if (w[0] + w[1] == 1) {
                    out << (w[0] ? "xm*" : "nxn*") << trig[w[2]];
                } else {
                    const size_t f = std::find(loop.factors.begin(), loop.factors.end(), std::make_pair(w[0], w[1])) - loop.factors.begin();
                    out << "mn" << id << "[" << st << "*m + " << 3 + f << "]*" << trig[w[2]];
                }
                out << ";" << std::endl;
            }
            for (size_t j = 0; j < loop.sets.size(); j++) {
                for (auto *m : loop.sets[j].members) {
                    const fourier_order o = fourier_order::unpack(m->num_cols);
                    const unsigned k = (o.b + o.c + (o.base ? 3u : 0u)) & 3u;
                    out << "                " << reg.at(m) << " = fma(p" << j << "_" << o.a << ", w" << o.b << "_" << o.c << "_" << k
                        << ", " << reg.at(m) << ");" << std::endl;
                    info.num_statements++;
                }
            }
            out << "            }" << std::endl;
            for (auto &set : loop.sets) {
                for (auto *m : set.members) {
                    const unsigned a = fourier_order::unpack(m->num_cols).a;
                    if (a == 2) out << "            " << reg.at(m) << " *= 2.0;" << std::endl;
                    if (a >= 3) out << "            " << reg.at(m) << " *= 6.0;" << std::endl;
                }
            }
            out << "        }" << std::endl;
        }

//  Every needed member of a spline family in one pass over the cell's coefficients (see graph::spline_2d):
//  per power of u the cubic in z and its derivatives by the Horner chain and its derivative recurrences,
//  then the same in u.  Members not asked for cost nothing.
        void emit_spline(const size_t g) {
            auto &grp = info.groups[g];
            if (grp.emitted) return;
            grp.emitted = true;
            using graph::op_t;
            const bool two = grp.op == op_t::spline_2d;
            const std::string id = std::to_string(g);
            const std::string x = emit(grp.arg0.get());
            const std::string z = two ? emit(grp.arg1.get()) : x;
            std::string cell;
            if (two) {
                cell = index_expr(x, grp.scale[0], grp.offset[0], grp.cells/grp.num_cols) + "*" + std::to_string(grp.num_cols) + "u + " +
                       index_expr(z, grp.scale[1], grp.offset[1], grp.num_cols);
            } else {
                cell = index_expr(x, grp.scale[0], grp.offset[0], grp.cells);
            }
            out << "        const double2 *sr" << id << " = reinterpret_cast<const double2 *> (" << group_pointer(g) << " + (" << cell << ")*"
                << grp.stride << "u);" << std::endl;
            info.num_statements++;
//  Which derivative orders along each direction are needed.
            std::set<unsigned> zorders;
            for (auto &o : grp.orders) zorders.insert(two ? o.second : o.first);
            auto need = [] (const std::set<unsigned> &set, std::initializer_list<unsigned> any) {
                for (unsigned v : any) if (set.count(v)) return true;
                return false;
            };
//  Chain + recurrences over coefficients k0..k3 (names) in variable `v`; returns the register of each order
//  WITHOUT the factors 2 and 6 of the second and third derivative.
            auto chain = [&] (const std::string &tag, const std::string &v, const std::array<std::string, 4> &k,
                              const std::set<unsigned> &orders) {
                std::array<std::string, 4> result;
                auto line = [&] (const std::string &name, const std::string &expr) {
                    out << "        const double " << name << " = " << expr << ";" << std::endl;
                    info.num_statements++;
                };
                if (need(orders, {0, 1, 2})) line(tag + "p1", "fma(" + v + ", " + k[3] + ", " + k[2] + ")");
                if (need(orders, {0, 1})) line(tag + "p2", "fma(" + v + ", " + tag + "p1, " + k[1] + ")");
                if (need(orders, {0})) { line(tag + "p3", "fma(" + v + ", " + tag + "p2, " + k[0] + ")"); result[0] = tag + "p3"; }
                if (need(orders, {1, 2})) line(tag + "q1", "fma(" + v + ", " + k[3] + ", " + tag + "p1)");
                if (need(orders, {1})) { line(tag + "q2", "fma(" + v + ", " + tag + "q1, " + tag + "p2)"); result[1] = tag + "q2"; }
                if (need(orders, {2})) { line(tag + "s1", "fma(" + v + ", " + k[3] + ", " + tag + "q1)"); result[2] = tag + "s1"; }
                if (need(orders, {3})) result[3] = k[3];
                return result;
            };
            const size_t pairs = grp.stride/2;
            for (size_t p = 0; p < pairs; p++) {
                out << "        const double2 sk" << id << "_" << p << " = " << (grp.staged ? "sr" + id + "[" + std::to_string(p) + "]"
                                                                                           : "__ldg(sr" + id + " + " + std::to_string(p) + ")")
                    << ";" << std::endl;
                info.num_statements++;
            }
            auto coefficient = [&] (const size_t index) {
                return "sk" + id + "_" + std::to_string(index/2) + (index & 1 ? ".y" : ".x");
            };
            static const double factorial[4] = {1.0, 1.0, 2.0, 6.0};
            if (!two) {
                const auto r = chain("s" + id + "_", x, {coefficient(0), coefficient(1), coefficient(2), coefficient(3)}, zorders);
                for (const graph::leaf_node *node : grp.spline_nodes) {
                    if (reg.count(node)) continue;
                    const unsigned a = graph::spline_order::unpack(node->num_cols).a;
                    const std::string name = "t" + local_id(node);
                    out << "        const double " << name << " = " << (a >= 2 ? literal(factorial[a]) + "*" : std::string()) << r[a] << ";" << std::endl;
                    info.num_statements++;
                    reg.emplace(node, name);
                }
                return;
            }
//  u = (r - offset)/scale as a multiplication, like every division by a constant in this front end.
            const double inv = 1.0/grp.scale[0];
            out << "        const double su" << id << " = (" << x << " - " << literal(grp.offset[0]) << ")*" << literal(inv) << ";" << std::endl;
            info.num_statements++;
            std::array<std::array<std::string, 4>, 4> rows;      // rows[i][b]: d^b/dz^b of the cubic multiplying u^i
            for (size_t i = 0; i < 4; i++) {
                rows[i] = chain("s" + id + "_" + std::to_string(i), z,
                                {coefficient(4*i), coefficient(4*i + 1), coefficient(4*i + 2), coefficient(4*i + 3)}, zorders);
            }
            for (const unsigned b : zorders) {
                std::set<unsigned> uorders;
                for (auto &o : grp.orders) if (o.second == b) uorders.insert(o.first);
                const auto r = chain("s" + id + "_u" + std::to_string(b), "su" + id, {rows[0][b], rows[1][b], rows[2][b], rows[3][b]}, uorders);
                for (const graph::leaf_node *node : grp.spline_nodes) {
                    if (reg.count(node)) continue;
                    const graph::spline_order o = graph::spline_order::unpack(node->num_cols);
                    if (o.b != b) continue;
                    double factor = factorial[o.b]*factorial[o.a];
                    for (unsigned i = 0; i < o.a; i++) factor *= inv;
                    const std::string name = "t" + local_id(node);
                    out << "        const double " << name << " = " << (factor != 1.0 ? literal(factor) + "*" : std::string()) << r[o.a] << ";" << std::endl;
                    info.num_statements++;
                    reg.emplace(node, name);
                }
            }
        }

        const std::string &emit(const graph::leaf_node *n) {
            n = strip(n);
            auto found = reg.find(n);
            if (found != reg.end()) return found->second;
            using graph::op_t;
            std::string rhs;
            if (n->op == op_t::constant) {
                return reg.emplace(n, literal(n->value)).first->second;
            }
            if (opt.hoist_invariants && !in_invariants && n->op != op_t::variable && is_invariant(n)) {
                const std::string name = hoisted(n, false);
                return reg.emplace(n, name).first->second;
            }
            if (n->op == op_t::variable) {
                for (size_t i = 0; i < info.inputs.size(); i++)
                    if (info.inputs[i].get() == n)
                        return reg.emplace(n, "v[" + std::to_string(i) + "]").first->second;
            }
            if (n->op == op_t::fourier) {
                emit_fourier_loop(floops[floop_of.at(n)]);
                return reg.at(n);
            }
            if (n->is_spline()) {
                emit_spline(slot.at(n).first);
                return reg.at(n);
            }
            if (n->is_index()) {
                const size_t g = slot.at(n).first;
                const size_t length = info.groups[g].cells;
                const std::string x = emit(n->args[1].get());
                std::string expr;
                if (n->op == op_t::index_1d) {
                    expr = index_expr(x, n->scale[0], n->offset[0], length);
                } else {
                    const std::string y = emit(n->args[2].get());
                    expr = index_expr(x, n->scale[0], n->offset[0], length/n->num_cols) + "*" +
                           std::to_string(n->num_cols) + "u + " + index_expr(y, n->scale[1], n->offset[1], n->num_cols);
                }
                const std::string name = "t" + local_id(n);
                out << "        const double " << name << " = tg[" << g << "][" << expr << "];" << std::endl;
                info.num_statements++;
                return reg.emplace(n, name).first->second;
            }
            if (n->is_piecewise()) {
                const auto [g, m] = slot.at(n);
                auto &grp = info.groups[g];
                if (index_reg[g].empty()) {
                    const std::string x = emit(grp.arg0.get());
                    std::string expr;
                    if (grp.op == op_t::piecewise_1d) {
                        expr = index_expr(x, grp.scale[0], grp.offset[0], grp.cells);
                    } else {
                        const std::string y = emit(grp.arg1.get());
                        expr = index_expr(x, grp.scale[0], grp.offset[0], grp.cells/grp.num_cols) + "*" +
                               std::to_string(grp.num_cols) + "u + " +
                               index_expr(y, grp.scale[1], grp.offset[1], grp.num_cols);
                    }
                    index_reg[g] = "row" + std::to_string(g);
                    out << "        const double *" << index_reg[g] << " = ";
                    if (grp.staged) {
                        out << "reinterpret_cast<const double *> (gfb::smem + " << grp.smem_offset << ")";
                    } else {
                        out << "tg[" << g << "]";
                    }
                    out << " + (" << expr << ")*" << grp.stride << "u;" << std::endl;
                    info.num_statements++;
                }
                const std::string name = "c" + std::to_string(g) + "_" + std::to_string(m);
                if (grp.staged) {
                    out << "        const double " << name << " = " << index_reg[g] << "[" << m << "];" << std::endl;
                } else {
                    const size_t p = m/2;
                    const std::string pname = "p" + std::to_string(g) + "_" + std::to_string(p);
                    if (!pair_loaded[g][p]) {
                        pair_loaded[g][p] = true;
                        out << "        const double2 " << pname << " = __ldg(reinterpret_cast<const double2 *> ("
                            << index_reg[g] << ") + " << p << ");" << std::endl;
                    }
                    out << "        const double " << name << " = " << pname << (m & 1 ? ".y" : ".x") << ";" << std::endl;
                }
                info.num_statements++;
                return reg.emplace(n, name).first->second;
            }

            if (n->op == op_t::sqrt && opt.fast_division && !denominators.count(n)) {
//  Nobody divides by this root: no need for its reciprocal.
                const std::string arg = emit(n->args[0].get());
                const std::string name = "t" + local_id(n);
                out << "        const double " << name << " = gfb::sqrt_only(" << arg << ");" << std::endl;
                info.num_statements++;
                return reg.emplace(n, name).first->second;
            }
            if (n->op == op_t::sqrt && opt.fast_division) {
                const std::string arg = emit(n->args[0].get());
                const std::string qname = "q" + local_id(n);
                const std::string name = "t" + local_id(n);
                out << "        const double " << qname << " = gfb::rsqrt(" << arg << ");" << std::endl;
                out << "        const double " << name << " = gfb::sqrt_from_rsqrt(" << arg << ", " << qname << ");" << std::endl;
                info.num_statements += 2;
                inv_reg.emplace(n, qname);      // 1/sqrt(x) is free: divisions by this node multiply by q
                return reg.emplace(n, name).first->second;
            }
            if (n->op == op_t::sin || n->op == op_t::cos) {
                const graph::leaf_node *arg = strip(n->args[0].get());
                auto both = trig.find(arg);
                if (both != trig.end() && both->second.first && both->second.second) {
                    const std::string areg = emit(arg);
                    const std::string sname = "t" + local_id(both->second.first);
                    const std::string cname = "t" + local_id(both->second.second);
                    out << "        double " << sname << ", " << cname << ";" << std::endl
                        << "        sincos(" << areg << ", &" << sname << ", &" << cname << ");" << std::endl;
                    info.num_statements += 2;
                    reg.emplace(both->second.first, sname);
                    reg.emplace(both->second.second, cname);
                    return reg.at(n);
                }
            }
            std::vector<std::string> a;
            for (size_t i = 0, ie = n->num_args(); i < ie; i++) {
                if (n->op == op_t::div && i == 1) {
                    const graph::leaf_node *d = strip(n->args[1].get());
                    if ((opt.share_reciprocals && denominators[d] > 1) || opt.fast_division) {
                        auto inv = inv_reg.find(d);
                        if (inv == inv_reg.end() && opt.hoist_invariants && !in_invariants && !d->is_constant() && is_invariant(d)) {
                            const std::string name = hoisted(d, true);
                            inv = inv_reg.emplace(d, name).first;
                        }
                        if (inv == inv_reg.end()) {
                            const std::string dreg = emit(d);
                            inv = inv_reg.find(d);      // a sqrt denominator registers its rsqrt while being emitted
                        }
                        if (inv == inv_reg.end()) {
                            const std::string dreg = emit(d);
                            const std::string iname = "i" + local_id(d);
                            out << "        const double " << iname << " = "
                                << (opt.fast_division ? "gfb::rcp(" + dreg + ")" : "1.0/" + dreg) << ";" << std::endl;
                            info.num_statements++;
                            info.num_reciprocals++;
                            inv = inv_reg.emplace(d, iname).first;
                        }
                        a.push_back(inv->second);
                        continue;
                    }
                }
                a.push_back(emit(n->args[i].get()));
            }
            switch (n->op) {
                case op_t::add: rhs = a[0] + " + " + a[1]; break;
                case op_t::sub: rhs = a[0] + " - " + a[1]; break;
                case op_t::mul: rhs = a[0] + "*" + a[1]; break;
                case op_t::div:
                    if (inv_reg.count(strip(n->args[1].get()))) {
                        rhs = a[0] + "*" + a[1];
                    } else {
                        rhs = a[0] + "/" + a[1];
                        info.num_divides++;
                    }
                    break;
                case op_t::fma: rhs = "fma(" + a[0] + ", " + a[1] + ", " + a[2] + ")"; break;
                case op_t::sqrt: rhs = "sqrt(" + a[0] + ")"; break;
                case op_t::exp: rhs = "exp(" + a[0] + ")"; break;
                case op_t::log: rhs = "log(" + a[0] + ")"; break;
                case op_t::erfi: rhs = "gfb::erfi(" + a[0] + ")"; break;
                case op_t::nan_to_zero: rhs = "(" + a[0] + " == " + a[0] + " ? " + a[0] + " : 0.0)"; break;
                case op_t::pow: rhs = "pow(" + a[0] + ", " + a[1] + ")"; break;
                case op_t::sin: rhs = "sin(" + a[0] + ")"; break;
                case op_t::cos: rhs = "cos(" + a[0] + ")"; break;
                case op_t::atan: rhs = "atan2(" + a[1] + ", " + a[0] + ")"; break;
                default: rhs = "0.0"; break;
            }
            const std::string name = "t" + local_id(n);
            out << "        const double " << name << " = " << rhs << ";" << std::endl;
            info.num_statements++;
            return reg.emplace(n, name).first->second;
        }

    public:
        body_emitter(std::ostringstream &out, const emit_options &opt, kernel_info &info) :
        out(out), opt(opt), info(info) {}

//------------------------------------------------------------------------------
///  @param[in] results  expressions evaluated by body, r[j] = results[j]
///  @param[in] apply    (input index, result index): v[first] = r[second] after each body call
///  @param[in] store_r  (pointer slot, result index) written on exit
///  @param[in] evolved  rk kinds: input index of each evolved component
///  @param[in] time_idx rk kinds: input index of the time variable or -1
//------------------------------------------------------------------------------
        void run(const std::vector<graph::leaf_ptr> &results,
                 const std::vector<std::pair<size_t, size_t>> &apply,
                 const std::vector<std::pair<size_t, size_t>> &store_r,
                 const std::vector<size_t> &evolved,
                 const int time_idx) {
            for (auto &r : results) scan(r.get());
            finish_mode_tables();
            pack_groups();
            index_reg.assign(info.groups.size(), "");
            pair_loaded.clear();
            for (auto &g : info.groups) pair_loaded.emplace_back((g.stride + 1)/2, false);

            const size_t ni = info.inputs.size(), nr = results.size(), ng = info.groups.size();
            const size_t np = ni + info.outputs.size();
            const std::string k = info.name + "_k";
            size_t staged_bytes = 0;
            for (auto &g : info.groups) if (g.staged) staged_bytes += g.bytes();
            out << std::endl << "struct " << k << " {" << std::endl
                << "    static constexpr int NI = " << ni << ", NR = " << nr << ", NG = " << ng
                << ", NP = " << np << ", NE = " << std::max<size_t> (evolved.size(), 1) << ", TI = " << time_idx << ";" << std::endl
                << "    static constexpr unsigned SMEM_BYTES = " << info.smem_bytes << ", STAGED_BYTES = " << staged_bytes << ";" << std::endl;
            auto table_fn = [&] (const char *type, const char *name, auto value) {
                out << "    __device__ static constexpr " << type << " " << name << "(const int g) { return ";
                for (size_t g = 0; g < ng; g++) out << "g == " << g << " ? " << value(info.groups[g]) << " : ";
                out << "0; }" << std::endl;
            };
            table_fn("bool", "group_staged", [] (const table_group &g) { return g.staged ? "true" : "false"; });
            table_fn("unsigned", "group_offset", [] (const table_group &g) { return std::to_string(g.smem_offset) + "u"; });
            table_fn("unsigned", "group_bytes", [] (const table_group &g) { return std::to_string(g.bytes()) + "u"; });
//  Pointer slot of each group: own table buffers follow the outputs; an alias group is a kernel input.
            {
                std::vector<size_t> slots;
                size_t own = 0;
                for (auto &g : info.groups) slots.push_back(g.alias_input >= 0 ? static_cast<size_t> (g.alias_input) : np + own++);
                out << "    __device__ static constexpr int group_slot(const int g) { return ";
                for (size_t g = 0; g < ng; g++) out << "g == " << g << " ? " << slots[g] << " : ";
                out << "0; }" << std::endl;
            }
//  An input that is only the array of index nodes is not a per-ray value: no load, any length.
            info.input_loaded.assign(ni, true);
            for (auto &g : info.groups) {
                if (g.alias_input < 0) continue;
                const size_t j = static_cast<size_t> (g.alias_input);
                info.input_loaded[j] = visited.count(info.inputs[j].get()) != 0 || info.input_written[j];
            }
            out << "    __device__ static constexpr int ev(const int e) { return ";
            for (size_t e = 0; e < evolved.size(); e++) out << "e == " << e << " ? " << evolved[e] << " : ";
            out << "0; }" << std::endl;

            out << "    __device__ static __forceinline__ void load(double (&v)[NI + 1], const gfb_args &a, const unsigned long long i) {" << std::endl;
            for (size_t j = 0; j < ni; j++) {
                if (!info.input_loaded[j]) {
                    continue;       // reached through index nodes only: its length need not match the kernel's
                } else if (info.input_written[j]) {
                    out << "        v[" << j << "] = a.ptr[" << j << "][i];" << std::endl;
                } else {
                    out << "        v[" << j << "] = __ldg(a.ptr[" << j << "] + i);" << std::endl;
                }
            }
            out << "    }" << std::endl;
            out << "    __device__ static __forceinline__ void apply(double (&v)[NI + 1], const double (&r)[NR + 1]) {" << std::endl;
            for (auto &[vi, ri] : apply) out << "        v[" << vi << "] = r[" << ri << "];" << std::endl;
            out << "    }" << std::endl;
            out << "    __device__ static __forceinline__ void store(const double (&v)[NI + 1], const double (&r)[NR + 1], const gfb_args &a, const unsigned long long i) {" << std::endl;
            for (size_t j = 0; j < ni; j++)
                if (info.input_written[j]) out << "        a.ptr[" << j << "][i] = v[" << j << "];" << std::endl;
            for (auto &[pi, ri] : store_r) out << "        a.ptr[" << pi << "][i] = r[" << ri << "];" << std::endl;
            out << "    }" << std::endl;

//  The statements of body() and of invariants() are collected apart (emit() switches between them).
            const std::string head = out.str();
            out.str("");
            std::vector<std::string> regs;
            for (auto &r : results) regs.push_back(emit(r.get()));
            for (size_t j = 0; j < nr; j++) out << "        r[" << j << "] = " << regs[j] << ";" << std::endl;
            const std::string body_text = out.str();
            out.str(head);
            out.seekp(0, std::ios_base::end);
            out << "    static constexpr int NH = " << num_invariants << ";" << std::endl
                << "    __device__ static __forceinline__ void invariants(const double (&v)[NI + 1], double (&h)[NH + 1]) {" << std::endl
                << text_other << "    }" << std::endl
                << "    __device__ static __forceinline__ void body(const double (&v)[NI + 1], const double (&h)[NH + 1], double (&r)[NR + 1], const double *(&tg)[NG + 1]) {" << std::endl
                << body_text << "    }" << std::endl << "};" << std::endl;

            out << "extern \"C\" __global__ void __launch_bounds__(" << opt.block_size << ", "
                << (opt.min_blocks ? std::to_string(opt.min_blocks) : std::string("GFB_MIN_BLOCKS")) << ") "
                << info.name << "(const __grid_constant__ gfb_args a) {" << std::endl;
            switch (info.kind) {
                case kernel_kind::generic: out << "    gfb::generic_item<" << k << "> (a);"; break;
                case kernel_kind::newton: out << "    gfb::newton_item<" << k << "> (a);"; break;
                case kernel_kind::rk2: out << "    gfb::runge_kutta<" << k << ", gfb::rk2_tableau> (a);"; break;
                case kernel_kind::rk4: out << "    gfb::runge_kutta<" << k << ", gfb::rk4_tableau> (a);"; break;
            }
            out << std::endl << "}" << std::endl;
        }
    };

//------------------------------------------------------------------------------
///  Emit a reference-style work item (inputs, outputs, setters).
//------------------------------------------------------------------------------
    inline kernel_info emit_item(std::ostringstream &out, const emit_options &opt,
                                 const kernel_kind kind, const std::string &name,
                                 graph::input_nodes<> inputs, graph::output_nodes<> outputs,
                                 graph::map_nodes<> setters, const size_t size) {
        kernel_info info;
        info.name = name;
        info.kind = kind;
        info.inputs = inputs;
        info.size = size;
        info.smem_bytes = 0;
        info.num_statements = info.num_divides = info.num_reciprocals = 0;
        info.input_written.assign(inputs.size(), false);

        std::vector<graph::leaf_ptr> results;
        std::vector<std::pair<size_t, size_t>> apply, store_r;
        for (auto &[expr, var] : setters) {
            if (expr.get() == var.get()) continue;      // self assignment (cpu_context.hpp:522)
            const size_t vi = std::find(inputs.begin(), inputs.end(), var) - inputs.begin();
            if (vi == inputs.size()) {
                std::cerr << "Kernel " << name << ": setter target is not an input." << std::endl;
                std::exit(-1);
            }
            info.input_written[vi] = true;
            apply.push_back({vi, results.size()});
            results.push_back(expr);
        }
        for (auto &o : outputs) {
            if (graph::variable_cast(o).get() &&
                std::find(inputs.begin(), inputs.end(), o) != inputs.end()) continue;   // already a buffer
            store_r.push_back({inputs.size() + info.outputs.size(), results.size()});
            info.outputs.push_back(o);
            results.push_back(o);
        }
        body_emitter(out, opt, info).run(results, apply, store_r, {}, -1);
        return info;
    }

//------------------------------------------------------------------------------
///  Emit a staged Runge-Kutta kernel: `rates[e]` is d(evolved[e])/dt.
//------------------------------------------------------------------------------
    inline kernel_info emit_runge_kutta(std::ostringstream &out, const emit_options &opt,
                                        const kernel_kind kind, const std::string &name,
                                        graph::input_nodes<> inputs,
                                        std::vector<graph::leaf_ptr> evolved,
                                        std::vector<graph::leaf_ptr> rates,
                                        graph::leaf_ptr time, graph::leaf_ptr dt,
                                        graph::leaf_ptr residual, const size_t size) {
        kernel_info info;
        info.name = name;
        info.kind = kind;
        info.inputs = inputs;
        info.size = size;
        info.smem_bytes = 0;
        info.num_statements = info.num_divides = info.num_reciprocals = 0;
        info.input_written.assign(inputs.size(), false);
        std::vector<size_t> ev;
        for (auto &e : evolved) {
            const size_t vi = std::find(inputs.begin(), inputs.end(), e) - inputs.begin();
            assert(vi < inputs.size() && "Evolved variable must be an input.");
            info.input_written[vi] = true;
            ev.push_back(vi);
        }
        int ti = -1;
        if (time.get()) {
            ti = static_cast<int> (std::find(inputs.begin(), inputs.end(), time) - inputs.begin());
            info.input_written[ti] = true;
        }
        std::vector<graph::leaf_ptr> results = rates;
        results.push_back(residual);
        results.push_back(dt);
        info.outputs.push_back(residual);
        body_emitter(out, opt, info).run(results, {}, {{inputs.size(), rates.size()}}, ev, ti);
        return info;
    }
}

#endif /* gfb_graph_emit_hpp */
