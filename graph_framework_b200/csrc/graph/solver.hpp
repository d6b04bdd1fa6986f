//------------------------------------------------------------------------------
//  solver.hpp -- ray integrators.
//
//  Mirrors /root/reference/graph_framework/solver.hpp: solver_interface
//  (:120-535: init, compile, step, sync_host, sync_device, check_residual,
//  print), rk2 (:550-666) and rk4 (:677-870).
//
//  The reference builds every Runge-Kutta stage as a separate symbolic copy of
//  the dispersion interface on pseudo variables and emits one kernel holding all
//  four copies, launched once per step.  Here the default (`STAGED = true`)
//  keeps ONE right-hand-side body and lets the hand-written skeleton run the
//  stage loop and the step loop in registers.  `STAGED = false` reproduces the
//  reference's construction (one generic item with the stages unrolled in the
//  graph); it exists to cross-check the skeleton and pseudo-variable handling.
//------------------------------------------------------------------------------
#ifndef gfb_graph_solver_hpp
#define gfb_graph_solver_hpp

#include "dispersion.hpp"
#include "binning.hpp"

namespace solver {
    using graph::leaf_ptr;

///  Whether solvers keep their rays sorted by table cell while stepping (binning.hpp); per host
///  thread, on unless the environment says GFB_BIN_RAYS=0.
    inline bool &bin_rays() {
        static thread_local bool on = !(std::getenv("GFB_BIN_RAYS") && std::string(std::getenv("GFB_BIN_RAYS")) == "0");
        return on;
    }

    template<dispersion::function DISPERSION_FUNCTION>
    class solver_interface {
    protected:
        typedef typename DISPERSION_FUNCTION::base T;
        static constexpr bool SAFE_MATH = DISPERSION_FUNCTION::safe_math;

        leaf_ptr w, kx, ky, kz, x, y, z, t;
        dispersion::dispersion_interface<DISPERSION_FUNCTION> D;
        leaf_ptr kx_next, ky_next, kz_next, x_next, y_next, z_next, t_next;
        leaf_ptr residual;
        workflow::manager<T, SAFE_MATH> work;
        const size_t index;
        newton_mode init_mode;
        equilibrium::cell_grid table_grid;
        workflow::ray_order<T, SAFE_MATH> order;
        bool order_wanted = bin_rays();
        size_t order_period = 0;        ///< 0: from the cell size and the step

///  Step size when it is a compile-time constant (how often the ray order is worth checking).
        virtual double step_size() const { return 0.0; }
///  After work.compile(): rays of tabulated equilibria are kept sorted by table cell.
        void setup_order() {
            if (!table_grid.dims || !order_wanted) return;
            std::vector<leaf_ptr> sort_by = table_grid.dims == 1 ? std::vector<leaf_ptr> {x} : std::vector<leaf_ptr> {x, y, z};
            order.configure(work, table_grid, sort_by, inputs(), {residual}, t->size(),
                            order_period ? order_period : table_grid.drift_steps(step_size()));
        }

        graph::input_nodes<T, SAFE_MATH> inputs() {
//  Argument order of solver.hpp:304-314.
            return {t, w, x, y, z, kx, ky, kz};
        }

    public:
        typedef DISPERSION_FUNCTION dispersion_function;
        typedef typename DISPERSION_FUNCTION::base base;

        solver_interface(leaf_ptr w, leaf_ptr kx, leaf_ptr ky, leaf_ptr kz,
                         leaf_ptr x, leaf_ptr y, leaf_ptr z, leaf_ptr t,
                         equilibrium::shared<T, SAFE_MATH> &eq,
                         const std::string &filename="", const size_t num_rays=0, const size_t index=0) :
        w(w), kx(kx), ky(ky), kz(kz), x(x), y(y), z(z), t(t),
        D(w, kx, ky, kz, x, y, z, t, eq), work(index), index(index), init_mode(newton_mode::per_ray),
        table_grid(eq->get_cell_grid()) {
            (void)filename; (void)num_rays;     // trajectory files are outside this back end
        }
        virtual ~solver_interface() {}

///  Before compile(): keep the rays sorted by table cell while stepping (default: yes for
///  tabulated equilibria, see solver::bin_rays()); `period` overrides how often the order is checked.
        void set_ray_order(const bool on, const size_t period=0) {
            order_wanted = on;
            order_period = period;
        }
///  Choose between the device-resident per-ray Newton solve (default) and the
///  reference's host-driven ensemble-maximum loop.
        void set_newton_mode(const newton_mode m) { init_mode = m; }

///  Solve D = 0 for one state variable (solver.hpp:254-274).
        virtual leaf_ptr init(leaf_ptr var, const T tolerance = 1.0E-30, const size_t max_iterations = 1000) final {
            residual = D.solve(var, inputs(), index, tolerance, max_iterations, init_mode);
            return residual;
        }
        virtual leaf_ptr init() final {
            residual = D.get_residual();
            return residual;
        }

///  Reference construction: a generic item with setters (solver.hpp:303-349).
        virtual void compile() {
            if (!residual.get()) residual = D.get_residual();
            graph::map_nodes<T, SAFE_MATH> setters = {
                {kx_next, kx}, {ky_next, ky}, {kz_next, kz},
                {x_next, x}, {y_next, y}, {z_next, z}, {t_next, t}
            };
            work.add_item(inputs(), {residual}, setters, graph::shared_random_state<T, SAFE_MATH> (),
                          "solver_kernel", t->size());
            work.compile();
            setup_order();
        }

        void sync_device() {
            order.restore();
            for (auto v : inputs()) work.copy_to_device(v, v->data());
        }
        void sync_host() {
            for (auto v : inputs()) {
                if (order.active()) order.copy_to_host(v, v->data());
                else work.copy_to_host(v, v->data());
            }
        }
        void step() { step(1); }
///  n steps; consecutive launches are fused by the device layer.  With a sorted ray order the
///  block is cut where the order is due for a check.
        void step(const size_t n) {
            size_t left = n;
            while (left) {
                const size_t piece = order.prepare(left);
                for (size_t i = 0; i < piece; i++) work.run();
                order.advanced(piece);
                left -= piece;
            }
        }
        T check_residual(const size_t i) { order.restore(); return work.check_value(i, residual); }
        void print(const size_t i) { order.restore(); work.print(i, {t, residual, w, x, y, z, kx, ky, kz}); }
        void wait() { work.wait(); }
///  The reference writes a NetCDF record here (solver.hpp:418-424); this back
///  end only provides the synchronisation point.
        void write_step() { work.wait(); }

        leaf_ptr get_residual() { return residual; }
        workflow::manager<T, SAFE_MATH> &get_work() { return work; }
///  The ray order policy (binning.hpp).  Code that reads device buffers of the rays by index through
///  get_work() must call get_order().restore() first.
        workflow::ray_order<T, SAFE_MATH> &get_order() { return order; }
        dispersion::dispersion_interface<DISPERSION_FUNCTION> &get_dispersion() { return D; }
        std::vector<leaf_ptr> state() { return inputs(); }
    };

    template<class S>
    concept method = std::is_base_of<solver_interface<typename S::dispersion_function>, S>::value;

//------------------------------------------------------------------------------
///  Second order Runge-Kutta (Heun), solver.hpp:550-666.
//------------------------------------------------------------------------------
    template<dispersion::function DISPERSION_FUNCTION, bool STAGED=true>
    class rk2 : public solver_interface<DISPERSION_FUNCTION> {
    protected:
        typedef typename DISPERSION_FUNCTION::base T;
        static constexpr bool SAFE_MATH = DISPERSION_FUNCTION::safe_math;
        leaf_ptr dt;
        virtual double step_size() const { return dt->is_constant() ? std::abs(dt->value) : 0.0; }
    public:
        rk2(leaf_ptr w, leaf_ptr kx, leaf_ptr ky, leaf_ptr kz, leaf_ptr x, leaf_ptr y, leaf_ptr z, leaf_ptr t,
            leaf_ptr dt, equilibrium::shared<T, SAFE_MATH> &eq,
            const std::string &filename="", const size_t num_rays=0, const size_t index=0) :
        solver_interface<DISPERSION_FUNCTION> (w, kx, ky, kz, x, y, z, t, eq, filename, num_rays, index), dt(dt) {
            if constexpr (!STAGED) {
                auto kx1 = dt*this->D.get_dkxdt(), ky1 = dt*this->D.get_dkydt(), kz1 = dt*this->D.get_dkzdt();
                auto x1 = dt*this->D.get_dxdt(), y1 = dt*this->D.get_dydt(), z1 = dt*this->D.get_dzdt();
                dispersion::dispersion_interface<DISPERSION_FUNCTION> D2(this->w,
                    graph::pseudo_variable(this->kx + kx1), graph::pseudo_variable(this->ky + ky1),
                    graph::pseudo_variable(this->kz + kz1), graph::pseudo_variable(this->x + x1),
                    graph::pseudo_variable(this->y + y1), graph::pseudo_variable(this->z + z1),
                    graph::pseudo_variable(this->t + dt), eq);
                auto kx2 = dt*D2.get_dkxdt(), ky2 = dt*D2.get_dkydt(), kz2 = dt*D2.get_dkzdt();
                auto x2 = dt*D2.get_dxdt(), y2 = dt*D2.get_dydt(), z2 = dt*D2.get_dzdt();
                this->kx_next = this->kx + (kx1 + kx2)/2.0;
                this->ky_next = this->ky + (ky1 + ky2)/2.0;
                this->kz_next = this->kz + (kz1 + kz2)/2.0;
                this->x_next = this->x + (x1 + x2)/2.0;
                this->y_next = this->y + (y1 + y2)/2.0;
                this->z_next = this->z + (z1 + z2)/2.0;
                this->t_next = this->t + dt;
            }
        }
        virtual void compile() {
            if (!this->residual.get()) this->residual = this->D.get_residual();
            if constexpr (STAGED) {
                this->work.add_runge_kutta_item(2, this->inputs(),
                    {this->kx, this->ky, this->kz, this->x, this->y, this->z},
                    {this->D.get_dkxdt(), this->D.get_dkydt(), this->D.get_dkzdt(),
                     this->D.get_dxdt(), this->D.get_dydt(), this->D.get_dzdt()},
                    this->t, dt, this->residual, "solver_kernel", this->t->size());
                this->work.compile();
                this->setup_order();
            } else {
                solver_interface<DISPERSION_FUNCTION>::compile();
            }
        }
    };

//------------------------------------------------------------------------------
///  Fourth order Runge-Kutta, solver.hpp:677-870.
//------------------------------------------------------------------------------
    template<dispersion::function DISPERSION_FUNCTION, bool STAGED=true>
    class rk4 : public solver_interface<DISPERSION_FUNCTION> {
    protected:
        typedef typename DISPERSION_FUNCTION::base T;
        static constexpr bool SAFE_MATH = DISPERSION_FUNCTION::safe_math;
        leaf_ptr dt;
        virtual double step_size() const { return dt->is_constant() ? std::abs(dt->value) : 0.0; }
    public:
        rk4(leaf_ptr w, leaf_ptr kx, leaf_ptr ky, leaf_ptr kz, leaf_ptr x, leaf_ptr y, leaf_ptr z, leaf_ptr t,
            leaf_ptr dt, equilibrium::shared<T, SAFE_MATH> &eq,
            const std::string &filename="", const size_t num_rays=0, const size_t index=0) :
        solver_interface<DISPERSION_FUNCTION> (w, kx, ky, kz, x, y, z, t, eq, filename, num_rays, index), dt(dt) {
            if constexpr (!STAGED) {
                typedef dispersion::dispersion_interface<DISPERSION_FUNCTION> DI;
                auto stage = [this, &eq] (leaf_ptr kxa, leaf_ptr kya, leaf_ptr kza,
                                          leaf_ptr xa, leaf_ptr ya, leaf_ptr za, leaf_ptr ta) {
                    return DI(this->w, graph::pseudo_variable(kxa), graph::pseudo_variable(kya),
                              graph::pseudo_variable(kza), graph::pseudo_variable(xa),
                              graph::pseudo_variable(ya), graph::pseudo_variable(za),
                              graph::pseudo_variable(ta), eq);
                };
                auto kx1 = dt*this->D.get_dkxdt(), ky1 = dt*this->D.get_dkydt(), kz1 = dt*this->D.get_dkzdt();
                auto x1 = dt*this->D.get_dxdt(), y1 = dt*this->D.get_dydt(), z1 = dt*this->D.get_dzdt();
                auto t_sub = this->t + dt/2.0;
                DI D2 = stage(this->kx + kx1/2.0, this->ky + ky1/2.0, this->kz + kz1/2.0,
                              this->x + x1/2.0, this->y + y1/2.0, this->z + z1/2.0, t_sub);
                auto kx2 = dt*D2.get_dkxdt(), ky2 = dt*D2.get_dkydt(), kz2 = dt*D2.get_dkzdt();
                auto x2 = dt*D2.get_dxdt(), y2 = dt*D2.get_dydt(), z2 = dt*D2.get_dzdt();
                DI D3 = stage(this->kx + kx2/2.0, this->ky + ky2/2.0, this->kz + kz2/2.0,
                              this->x + x2/2.0, this->y + y2/2.0, this->z + z2/2.0, t_sub);
                auto kx3 = dt*D3.get_dkxdt(), ky3 = dt*D3.get_dkydt(), kz3 = dt*D3.get_dkzdt();
                auto x3 = dt*D3.get_dxdt(), y3 = dt*D3.get_dydt(), z3 = dt*D3.get_dzdt();
                this->t_next = this->t + dt;
                DI D4 = stage(this->kx + kx3, this->ky + ky3, this->kz + kz3,
                              this->x + x3, this->y + y3, this->z + z3, this->t_next);
                auto kx4 = dt*D4.get_dkxdt(), ky4 = dt*D4.get_dkydt(), kz4 = dt*D4.get_dkzdt();
                auto x4 = dt*D4.get_dxdt(), y4 = dt*D4.get_dydt(), z4 = dt*D4.get_dzdt();
                this->kx_next = this->kx + (kx1 + 2.0*(kx2 + kx3) + kx4)/6.0;
                this->ky_next = this->ky + (ky1 + 2.0*(ky2 + ky3) + ky4)/6.0;
                this->kz_next = this->kz + (kz1 + 2.0*(kz2 + kz3) + kz4)/6.0;
                this->x_next = this->x + (x1 + 2.0*(x2 + x3) + x4)/6.0;
                this->y_next = this->y + (y1 + 2.0*(y2 + y3) + y4)/6.0;
                this->z_next = this->z + (z1 + 2.0*(z2 + z3) + z4)/6.0;
            }
        }
        virtual void compile() {
            if (!this->residual.get()) this->residual = this->D.get_residual();
            if constexpr (STAGED) {
                this->work.add_runge_kutta_item(4, this->inputs(),
                    {this->kx, this->ky, this->kz, this->x, this->y, this->z},
                    {this->D.get_dkxdt(), this->D.get_dkydt(), this->D.get_dkzdt(),
                     this->D.get_dxdt(), this->D.get_dydt(), this->D.get_dzdt()},
                    this->t, dt, this->residual, "solver_kernel", this->t->size());
                this->work.compile();
                this->setup_order();
            } else {
                solver_interface<DISPERSION_FUNCTION>::compile();
            }
        }
    };

//------------------------------------------------------------------------------
///  Adaptive time step RK4, solver.hpp:881-1006: before every step a two-unknown Newton solve
///  on (dt, lambda) of  1/dt + lambda D(next state)^2  picks the step length, then the RK4 step is
///  taken with that dt.  dt is a per-ray VARIABLE here.  Built on the graph-unrolled rk4 (the
///  loss needs the symbolic next state).  The Newton solve runs per ray on the device unless
///  set_newton_mode(ensemble) asks for the reference's host-driven loop.
//------------------------------------------------------------------------------
    template<dispersion::function DISPERSION_FUNCTION>
    class adaptive_rk4 : public rk4<DISPERSION_FUNCTION, false> {
    protected:
        typedef typename DISPERSION_FUNCTION::base T;
        static constexpr bool SAFE_MATH = DISPERSION_FUNCTION::safe_math;
        dispersion::dispersion_interface<DISPERSION_FUNCTION> D_next;
        leaf_ptr dt_var;
    public:
        adaptive_rk4(leaf_ptr w, leaf_ptr kx, leaf_ptr ky, leaf_ptr kz, leaf_ptr x, leaf_ptr y, leaf_ptr z, leaf_ptr t,
                     leaf_ptr dt, equilibrium::shared<T, SAFE_MATH> &eq,
                     const std::string &filename="", const size_t num_rays=0, const size_t index=0) :
        rk4<DISPERSION_FUNCTION, false> (w, kx, ky, kz, x, y, z, t, dt, eq, filename, num_rays, index),
        D_next(w, graph::pseudo_variable(this->kx_next), graph::pseudo_variable(this->ky_next),
               graph::pseudo_variable(this->kz_next), graph::pseudo_variable(this->x_next),
               graph::pseudo_variable(this->y_next), graph::pseudo_variable(this->z_next),
               graph::pseudo_variable(this->t_next), eq),
        dt_var(dt) {
            assert(graph::variable_cast(dt).get() && "adaptive_rk4 needs dt to be a variable.");
        }
        virtual void compile() final {
            if (!this->residual.get()) this->residual = this->D.get_residual();
            auto lambda = graph::variable(dt_var->size(), 1.0, "\\lambda");
            auto d_next = D_next.get_d()->remove_pseudo();
            auto loss = graph::one()/dt_var + lambda*d_next*d_next;
            auto inputs = this->inputs();
            inputs.push_back(dt_var);
            auto newton_inputs = inputs;
            newton_inputs.push_back(lambda);
            solver::newton<T, SAFE_MATH> (this->work, {dt_var, lambda}, newton_inputs, loss,
                                          graph::shared_random_state<T, SAFE_MATH> (), 1.0E-30, 1000, 1.0, this->init_mode);
            graph::map_nodes<T, SAFE_MATH> setters = {
                {this->kx_next, this->kx}, {this->ky_next, this->ky}, {this->kz_next, this->kz},
                {this->x_next, this->x}, {this->y_next, this->y}, {this->z_next, this->z}, {this->t_next, this->t}
            };
            this->work.add_item(inputs, {this->residual}, setters, graph::shared_random_state<T, SAFE_MATH> (),
                                "solver_kernel", this->t->size());
            this->work.compile();               // rays stay in the caller's order: dt and lambda are per-ray state too
        }
        leaf_ptr get_dt() { return dt_var; }
    };

//------------------------------------------------------------------------------
///  Second order symplectic split (position half step, momentum step, position half step) for
///  separable Hamiltonians, solver.hpp:1016-1130.  Built in the graph like the reference does
///  (stage arguments are pseudo variables) and run as a generic fused item.
//------------------------------------------------------------------------------
    template<dispersion::function DISPERSION_FUNCTION>
    class split_simplextic : public solver_interface<DISPERSION_FUNCTION> {
    protected:
        typedef typename DISPERSION_FUNCTION::base T;
        static constexpr bool SAFE_MATH = DISPERSION_FUNCTION::safe_math;
    public:
        split_simplextic(leaf_ptr w, leaf_ptr kx, leaf_ptr ky, leaf_ptr kz, leaf_ptr x, leaf_ptr y, leaf_ptr z,
                         leaf_ptr t, leaf_ptr dt, equilibrium::shared<T, SAFE_MATH> &eq,
                         const std::string &filename="", const size_t num_rays=0, const size_t index=0) :
        solver_interface<DISPERSION_FUNCTION> (w, kx, ky, kz, x, y, z, t, eq, filename, num_rays, index) {
            typedef dispersion::dispersion_interface<DISPERSION_FUNCTION> DI;
//  Separable: dk/dt must not depend on k, dx/dt must not depend on x (solver.hpp:1060-1079).
            bool separable = true;
            for (auto rate : {this->D.get_dkxdt(), this->D.get_dkydt(), this->D.get_dkzdt()})
                for (auto v : {kx, ky, kz}) separable = separable && rate->df(v)->is_constant(0.0);
            for (auto rate : {this->D.get_dxdt(), this->D.get_dydt(), this->D.get_dzdt()})
                for (auto v : {x, y, z}) separable = separable && rate->df(v)->is_constant(0.0);
            if (!separable) {
                std::cerr << "split_simplextic: Hamiltonian is not separable." << std::endl;
                std::abort();
            }
            this->t_next = t + dt;
            auto x1 = x + dt*this->D.get_dxdt()/2.0;
            auto y1 = y + dt*this->D.get_dydt()/2.0;
            auto z1 = z + dt*this->D.get_dzdt()/2.0;
            auto pv = [] (leaf_ptr a) { return graph::pseudo_variable(a); };
            DI D2(w, pv(kx), pv(ky), pv(kz), pv(x1), pv(y1), pv(z1), pv(t), eq);
            this->kx_next = kx + dt*D2.get_dkxdt();
            this->ky_next = ky + dt*D2.get_dkydt();
            this->kz_next = kz + dt*D2.get_dkzdt();
            DI D3(w, pv(this->kx_next), pv(this->ky_next), pv(this->kz_next), pv(x1), pv(y1), pv(z1), pv(t), eq);
            this->x_next = x1 + dt*D3.get_dxdt()/2.0;
            this->y_next = y1 + dt*D3.get_dydt()/2.0;
            this->z_next = z1 + dt*D3.get_dzdt()/2.0;
        }
    };
}

#endif /* gfb_graph_solver_hpp */
