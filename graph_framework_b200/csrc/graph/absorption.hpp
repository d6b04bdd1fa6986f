//------------------------------------------------------------------------------
//  absorption.hpp -- power absorption along traced rays.
//
//  Mirrors /root/reference/graph_framework/absorption.hpp (weak_damping :327-484) and the power
//  stage of the reference driver (/root/reference/graph_driver/xrays.cpp:674-793, bin_power).
//
//  In the reference these are two further passes over the trajectory FILES: every record is read
//  back from NetCDF, copied to the device, one kernel is run, and the result is written back.  Here
//  both kernels can be attached to the solver's own workflow manager: they then read the ray state
//  where the Runge-Kutta kernel left it in HBM and run between two blocks of steps, so a trace with
//  absorption costs two small launches per record and no transfer.  The stand-alone constructors
//  (own manager, explicit copy_to_device) keep the reference's calling sequence.
//
//  kamp is complex in the reference; only its imaginary part feeds the power stage
//  (reference_imag_variable, xrays.cpp:743).  Both parts are produced here as two real variables.
//------------------------------------------------------------------------------
#ifndef gfb_graph_absorption_hpp
#define gfb_graph_absorption_hpp

#include "dispersion.hpp"

namespace absorption {
    using graph::leaf_ptr;

    template<typename T=double, bool SAFE_MATH=false>
    class method {
    public:
        typedef T base;
        static constexpr bool safe_math = SAFE_MATH;
        virtual ~method() {}
        virtual void compile() = 0;
        virtual void run(const size_t time_index) = 0;
    };

//------------------------------------------------------------------------------
///  Weak damping approximation: k_amp = |k| - D_warm/(k_hat . dD_cold/dk)
///  (absorption.hpp:395-412).
//------------------------------------------------------------------------------
    template<typename T=double, bool SAFE_MATH=false>
    class weak_damping final : public method<T, SAFE_MATH> {
    private:
        leaf_ptr kamp_re, kamp_im, w, kx, ky, kz, x, y, z, t;
        leaf_ptr re_next, im_next;
        std::unique_ptr<workflow::manager<T, SAFE_MATH>> own;
        workflow::manager<T, SAFE_MATH> &work;
        size_t item;

        void build(equilibrium::shared<T, SAFE_MATH> &eq) {
            auto k_vec = kx*eq->get_esup1(x, y, z) + ky*eq->get_esup2(x, y, z) + kz*eq->get_esup3(x, y, z);
            auto k_unit = k_vec->unit();
            auto Dc = dispersion::cold_plasma_expansion<T, SAFE_MATH> ().D(w, k_vec, x, y, z, t, eq);
            auto Dw = dispersion::hot_plasma_expansion<T, dispersion::z_erfi<T, SAFE_MATH>, SAFE_MATH> ()
                          .D_complex(w, k_vec, x, y, z, t, eq);
            std::vector<leaf_ptr> grad = dispersion::reverse_mode() ? graph::gradient(Dc, {kx, ky, kz}) :
                                         std::vector<leaf_ptr> {Dc->df(kx), Dc->df(ky), Dc->df(kz)};
            auto slope = k_unit->dot(grad[0]*eq->get_esup1(x, y, z) + grad[1]*eq->get_esup2(x, y, z) +
                                     grad[2]*eq->get_esup3(x, y, z));
            re_next = k_vec->length() - Dw.re/slope;
            im_next = -1.0*(Dw.im/slope);
//  Argument order of absorption.hpp:414-424 with kamp split in two.
            graph::input_nodes<T, SAFE_MATH> inputs = {kamp_re, kamp_im, kx, ky, kz, x, y, z, t, w};
            graph::map_nodes<T, SAFE_MATH> setters = {{re_next, kamp_re}, {im_next, kamp_im}};
            item = work.add_side_item(inputs, {}, setters, graph::shared_random_state<T, SAFE_MATH> (),
                                      "weak_damping_kimg_kernel", w->size());
        }

    public:
///  Reference calling sequence: own device context `index`; the caller copies a record to the
///  device (copy_to_device) and calls run().
        weak_damping(leaf_ptr kamp_re, leaf_ptr kamp_im, leaf_ptr w, leaf_ptr kx, leaf_ptr ky, leaf_ptr kz,
                     leaf_ptr x, leaf_ptr y, leaf_ptr z, leaf_ptr t,
                     equilibrium::shared<T, SAFE_MATH> &eq, const std::string &filename="", const size_t index=0) :
        kamp_re(kamp_re), kamp_im(kamp_im), w(w), kx(kx), ky(ky), kz(kz), x(x), y(y), z(z), t(t),
        own(std::make_unique<workflow::manager<T, SAFE_MATH>> (index)), work(*own) {
            (void)filename;
            build(eq);
        }
///  Attached: the kernel joins `manager` (normally solver.get_work()) BEFORE it is compiled and
///  shares its device buffers.
        weak_damping(workflow::manager<T, SAFE_MATH> &manager, leaf_ptr kamp_re, leaf_ptr kamp_im,
                     leaf_ptr w, leaf_ptr kx, leaf_ptr ky, leaf_ptr kz, leaf_ptr x, leaf_ptr y, leaf_ptr z, leaf_ptr t,
                     equilibrium::shared<T, SAFE_MATH> &eq) :
        kamp_re(kamp_re), kamp_im(kamp_im), w(w), kx(kx), ky(ky), kz(kz), x(x), y(y), z(z), t(t), work(manager) {
            build(eq);
        }

        virtual void compile() { if (own) work.compile(); }
///  One evaluation at the state currently on the device.
        virtual void run(const size_t time_index=0) { (void)time_index; work.run_side(item); }
        void sync_device() {
            for (auto v : {w, kx, ky, kz, x, y, z, t}) work.copy_to_device(v, v->data());
        }
        void sync_host() {
            work.copy_to_host(kamp_re, kamp_re->data());
            work.copy_to_host(kamp_im, kamp_im->data());
        }
        void wait() { work.wait(); }
        workflow::manager<T, SAFE_MATH> &get_work() { return work; }
///  The expressions the kernel assigns to (Re, Im) k_amp, e.g. for host evaluation.
        leaf_ptr get_real_expression() { return re_next; }
        leaf_ptr get_imaginary_expression() { return im_next; }
    };

//------------------------------------------------------------------------------
///  The power stage (xrays.cpp:693-736): path length since the previous record, running
///  optical depth k_sum, transmitted power and the power lost in the segment.
///      dl = |X - X_last|, p_next = exp(-2 k_sum), d_power = |p_next - power|,
///      k_sum += kamp dl, power = p_next, X_last = X.
///  As in the reference p_next uses k_sum BEFORE the current segment is added.
//------------------------------------------------------------------------------
    template<typename T=double, bool SAFE_MATH=false>
    class power_item {
    private:
        leaf_ptr x, y, z, x_last, y_last, z_last, kamp, power, k_sum, d_power;
        std::unique_ptr<workflow::manager<T, SAFE_MATH>> own;
        workflow::manager<T, SAFE_MATH> &work;
        size_t item;

        void build(equilibrium::shared<T, SAFE_MATH> &eq) {
            auto dl = graph::vector(eq->get_x(x, y, z) - eq->get_x(x_last, y_last, z_last),
                                    eq->get_y(x, y, z) - eq->get_y(x_last, y_last, z_last),
                                    eq->get_z(x, y, z) - eq->get_z(x_last, y_last, z_last))->length();
            auto k_next = kamp*dl + k_sum;
            auto p_next = graph::exp(-2.0*k_sum);
            auto difference = p_next - power;
            d_power = graph::sqrt(difference*difference);
            item = work.add_side_item({x, y, z, x_last, y_last, z_last, kamp, power, k_sum}, {d_power},
                                      {{x, x_last}, {y, y_last}, {z, z_last}, {p_next, power}, {k_next, k_sum}},
                                      graph::shared_random_state<T, SAFE_MATH> (), "power", x->size());
        }

    public:
        power_item(leaf_ptr x, leaf_ptr y, leaf_ptr z, leaf_ptr x_last, leaf_ptr y_last, leaf_ptr z_last,
                   leaf_ptr kamp, leaf_ptr power, leaf_ptr k_sum,
                   equilibrium::shared<T, SAFE_MATH> &eq, const size_t index=0) :
        x(x), y(y), z(z), x_last(x_last), y_last(y_last), z_last(z_last), kamp(kamp), power(power), k_sum(k_sum),
        own(std::make_unique<workflow::manager<T, SAFE_MATH>> (index)), work(*own) {
            build(eq);
        }
        power_item(workflow::manager<T, SAFE_MATH> &manager,
                   leaf_ptr x, leaf_ptr y, leaf_ptr z, leaf_ptr x_last, leaf_ptr y_last, leaf_ptr z_last,
                   leaf_ptr kamp, leaf_ptr power, leaf_ptr k_sum, equilibrium::shared<T, SAFE_MATH> &eq) :
        x(x), y(y), z(z), x_last(x_last), y_last(y_last), z_last(z_last), kamp(kamp), power(power), k_sum(k_sum),
        work(manager) {
            build(eq);
        }
        void compile() { if (own) work.compile(); }
        void run() { work.run_side(item); }
        leaf_ptr get_d_power() { return d_power; }
        workflow::manager<T, SAFE_MATH> &get_work() { return work; }
    };
}

#endif /* gfb_graph_absorption_hpp */
