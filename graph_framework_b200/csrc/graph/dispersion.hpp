//------------------------------------------------------------------------------
//  dispersion.hpp -- dispersion functions D(w, k, x, t) and the ray equations.
//
//  Mirrors /root/reference/graph_framework/dispersion.hpp: the
//  dispersion_function interface (:160-189), simple (:449-477), physics
//  constants (:486-502), bohm_gross (:510-575), light_wave (:583-625),
//  acoustic_wave (:633-690), gaussian_well (:698-731), ion_cyclotron (:739-783),
//  ordinary_wave (:785-829), extra_ordinary_wave (:838-895), cold_plasma
//  (:903-1008) and dispersion_interface (:1321-1634) which derives
//      dx/dt = -dD/dk / dD/dw,   dk/dt = (dD/dx - dD/dk_vec . dk_vec/dx) / dD/dw
//  by symbolic differentiation.
//------------------------------------------------------------------------------
#ifndef gfb_graph_dispersion_hpp
#define gfb_graph_dispersion_hpp

#include "equilibrium.hpp"

namespace dispersion {
    using graph::leaf_ptr;
    using graph::vector_ptr;

///  w_p^2/c^2 for a species (dispersion.hpp:108-117).
    inline leaf_ptr build_plasma_frequency(leaf_ptr n, const double q, const double m, const double c,
                                           const double epsilon0) {
        return n*q*q/(epsilon0*m*c*c);
    }
///  w_c/c for a species (dispersion.hpp:132-139).
    inline leaf_ptr build_cyclotron_frequency(const double q, leaf_ptr b, const double m, const double c) {
        return q*b/(m*c);
    }

    template<typename T=double, bool SAFE_MATH=false>
    class dispersion_function {
    public:
        virtual ~dispersion_function() {}
        virtual leaf_ptr D(leaf_ptr w, vector_ptr k_vec, leaf_ptr x, leaf_ptr y, leaf_ptr z, leaf_ptr t,
                           equilibrium::shared<T, SAFE_MATH> &eq) = 0;
///  Reference compatibility (see reference_defects()): the term the reference's symbolic
///  dD/d(coordinate `axis`) carries on top of the true derivative; null when there is none.
        virtual leaf_ptr reference_defect(leaf_ptr, vector_ptr, leaf_ptr, leaf_ptr, leaf_ptr, leaf_ptr,
                                          equilibrium::shared<T, SAFE_MATH> &, int &axis) {
            axis = -1;
            return leaf_ptr();
        }
        typedef T base;
        static constexpr bool safe_math = SAFE_MATH;
    };

///  D = (1000 (x - e^-t) - e^-t) kx + w  (dispersion.hpp:197-229).
    template<typename T=double, bool SAFE_MATH=false>
    class stiff final : public dispersion_function<T, SAFE_MATH> {
    public:
        virtual leaf_ptr D(leaf_ptr w, vector_ptr k_vec, leaf_ptr x, leaf_ptr, leaf_ptr, leaf_ptr t,
                           equilibrium::shared<T, SAFE_MATH> &) {
            return (1.0E3*(x - graph::exp(-t)) - graph::exp(-t))*k_vec->get_x() + w;
        }
    };

///  D = n_par^2 + n_perp^2 - 1 with c = 1.
    template<typename T=double, bool SAFE_MATH=false>
    class simple final : public dispersion_function<T, SAFE_MATH> {
    public:
        virtual leaf_ptr D(leaf_ptr w, vector_ptr k_vec, leaf_ptr, leaf_ptr, leaf_ptr, leaf_ptr,
                           equilibrium::shared<T, SAFE_MATH> &) {
            const T c = 1.0;
            auto npar2 = k_vec->get_z()*k_vec->get_z()*c*c/(w*w);
            auto nperp2 = (k_vec->get_x()*k_vec->get_x() + k_vec->get_y()*k_vec->get_y())*c*c/(w*w);
            return npar2 + nperp2 - c;
        }
    };

    template<typename T=double, bool SAFE_MATH=false>
    class physics : public dispersion_function<T, SAFE_MATH> {
    protected:
        const T epsilon0 = 8.8541878138E-12;
        const T mu0 = M_PI*4.0E-7;
        const T q = 1.602176634E-19;
        const T me = 9.1093837015E-31;
        const T c = static_cast<T> (1.0)/std::sqrt(epsilon0*mu0);

///  k_par^2 with the reference's "no field -> |k|^2" rule (dispersion.hpp:552-560).
        leaf_ptr kpara2(vector_ptr b_vec, vector_ptr k_vec) {
            if (b_vec->length()->is_match(graph::zero())) return k_vec->dot(k_vec);
            auto kpara = b_vec->unit()->dot(k_vec);
            return kpara*kpara;
        }
    };

///  D = w_pe^2 + 3/2 k_par^2 v_th^2 - w^2.
    template<typename T=double, bool SAFE_MATH=false>
    class bohm_gross final : public physics<T, SAFE_MATH> {
    public:
        virtual leaf_ptr D(leaf_ptr w, vector_ptr k_vec, leaf_ptr x, leaf_ptr y, leaf_ptr z, leaf_ptr,
                           equilibrium::shared<T, SAFE_MATH> &eq) {
            auto wpe2 = build_plasma_frequency(eq->get_electron_density(x, y, z), this->q, this->me, this->c, this->epsilon0);
            auto te = eq->get_electron_temperature(x, y, z);
            auto vterm2 = static_cast<T> (2.0)*this->q*te/(this->me*this->c*this->c);
            return wpe2 + 3.0/2.0*this->kpara2(eq->get_magnetic_field(x, y, z), k_vec)*vterm2 - w*w;
        }
    };

///  D = w_pe^2 + k^2 - w^2.
    template<typename T=double, bool SAFE_MATH=false>
    class light_wave final : public physics<T, SAFE_MATH> {
    public:
        virtual leaf_ptr D(leaf_ptr w, vector_ptr k_vec, leaf_ptr x, leaf_ptr y, leaf_ptr z, leaf_ptr,
                           equilibrium::shared<T, SAFE_MATH> &eq) {
            auto wpe2 = build_plasma_frequency(eq->get_electron_density(x, y, z), this->q, this->me, this->c, this->epsilon0);
            assert(eq->get_magnetic_field(x, y, z)->length()->is_match(graph::zero()) &&
                   "Expected equilibrium with no magnetic field.");
            return wpe2 + k_vec->dot(k_vec) - w*w;
        }
    };

///  D = k_par^2 v_s^2 - w^2.
    template<typename T=double, bool SAFE_MATH=false>
    class acoustic_wave final : public physics<T, SAFE_MATH> {
    public:
        virtual leaf_ptr D(leaf_ptr w, vector_ptr k_vec, leaf_ptr x, leaf_ptr y, leaf_ptr z, leaf_ptr,
                           equilibrium::shared<T, SAFE_MATH> &eq) {
            const T mi = eq->get_ion_mass(0);
            auto te = eq->get_electron_temperature(x, y, z);
            auto ti = eq->get_ion_temperature(0, x, y, z);
            const T gamma = 3.0;
            auto vs2 = (this->q*te + gamma*this->q*ti)/(mi*this->c*this->c);
            return this->kpara2(eq->get_magnetic_field(x, y, z), k_vec)*vs2 - w*w;
        }
    };

///  D = n_par^2 + n_perp^2 - (1 - exp(-(x^2 + y^2)/0.1)/2).
    template<typename T=double, bool SAFE_MATH=false>
    class gaussian_well final : public dispersion_function<T, SAFE_MATH> {
    public:
        virtual leaf_ptr D(leaf_ptr w, vector_ptr k_vec, leaf_ptr x, leaf_ptr y, leaf_ptr, leaf_ptr,
                           equilibrium::shared<T, SAFE_MATH> &) {
            const T c = 1.0;
            auto well = c - 0.5*graph::exp(-(x*x + y*y)/0.1);
            auto npar2 = k_vec->get_z()*k_vec->get_z()*c*c/(w*w);
            auto nperp2 = (k_vec->get_x()*k_vec->get_x() + k_vec->get_y()*k_vec->get_y())*c*c/(w*w);
            return npar2 + nperp2 - well;
        }
    };

///  D = w_ce - k_perp^2 v_s^2 - w^2  (as written in the reference, dispersion.hpp:739-783).
    template<typename T=double, bool SAFE_MATH=false>
    class ion_cyclotron final : public physics<T, SAFE_MATH> {
    public:
        virtual leaf_ptr D(leaf_ptr w, vector_ptr k_vec, leaf_ptr x, leaf_ptr y, leaf_ptr z, leaf_ptr,
                           equilibrium::shared<T, SAFE_MATH> &eq) {
            const T mi = eq->get_ion_mass(0);
            auto te = eq->get_electron_temperature(x, y, z);
            auto ti = eq->get_ion_temperature(0, x, y, z);
            const T gamma = 3.0;
            auto vs2 = (this->q*te + gamma*this->q*ti)/(mi*this->c*this->c);
            auto b_vec = eq->get_magnetic_field(x, y, z);
            auto wce = build_cyclotron_frequency(-this->q, b_vec->length(), this->me, this->c);
            auto kperp = b_vec->unit()->cross(k_vec)->length();
            return wce - kperp*kperp*vs2 - w*w;
        }
    };

///  O-mode: D = 1 - w_pe^2/w^2 - n_perp^2.
    template<typename T=double, bool SAFE_MATH=false>
    class ordinary_wave final : public physics<T, SAFE_MATH> {
    public:
        virtual leaf_ptr D(leaf_ptr w, vector_ptr k_vec, leaf_ptr x, leaf_ptr y, leaf_ptr z, leaf_ptr,
                           equilibrium::shared<T, SAFE_MATH> &eq) {
            auto wpe2 = build_plasma_frequency(eq->get_electron_density(x, y, z), this->q, this->me, this->c, this->epsilon0);
            auto n = k_vec/w;
            auto b_hat = eq->get_magnetic_field(x, y, z)->unit();
            auto nperp = b_hat->cross(n);
            auto nperp2 = nperp->dot(nperp);
            return 1.0 - wpe2/(w*w) - nperp2;
        }
    };

///  X-mode: D = 1 - w_pe^2/w^2 (w^2 - w_pe^2)/(w^2 - w_h^2) - n_perp^2.
    template<typename T=double, bool SAFE_MATH=false>
    class extra_ordinary_wave final : public physics<T, SAFE_MATH> {
    public:
        virtual leaf_ptr D(leaf_ptr w, vector_ptr k_vec, leaf_ptr x, leaf_ptr y, leaf_ptr z, leaf_ptr,
                           equilibrium::shared<T, SAFE_MATH> &eq) {
            auto wpe2 = build_plasma_frequency(eq->get_electron_density(x, y, z), this->q, this->me, this->c, this->epsilon0);
            auto b_vec = eq->get_magnetic_field(x, y, z);
            auto wec = build_cyclotron_frequency(-this->q, b_vec->length(), this->me, this->c);
            auto n = k_vec/w;
            auto nperp = b_vec->unit()->cross(n);
            auto nperp2 = nperp->dot(nperp);
            auto wh = wpe2 + wec*wec;
            auto w2 = w*w;
            return 1.0 - wpe2/w2*(w2 - wpe2)/(w2 - wh) - nperp2;
        }
    };

///  Cold plasma determinant with electrons and every ion species of the equilibrium.
    template<typename T=double, bool SAFE_MATH=false>
    class cold_plasma : public physics<T, SAFE_MATH> {
    protected:
///  The named sub-expressions of the determinant (nodes are interned: building them twice costs nothing).
        struct elements {
            leaf_ptr w2, b_sq, b_len, npara2, nperp2, m11, m12, m13, m22, m33;
        };
        elements build(leaf_ptr w, vector_ptr k_vec, leaf_ptr x, leaf_ptr y, leaf_ptr z,
                       equilibrium::shared<T, SAFE_MATH> &eq) {
            auto wpe2 = build_plasma_frequency(eq->get_electron_density(x, y, z), this->q, this->me, this->c, this->epsilon0);
            auto b_vec = eq->get_magnetic_field(x, y, z);
            auto b_len = b_vec->length();
            auto ec = build_cyclotron_frequency(-this->q, b_len, this->me, this->c);

            auto w2 = w*w;
            auto denome = 1.0 - ec*ec/w2;
            auto e11 = 1.0 - (wpe2/w2)/denome;
            auto e12 = ((ec/w)*(wpe2/w2))/denome;
            auto e33 = wpe2;
            for (size_t i = 0, ie = eq->get_num_ion_species(); i < ie; i++) {
                const T mi = eq->get_ion_mass(i);
                const T charge = static_cast<T> (eq->get_ion_charge(i))*this->q;
                auto wpi2 = build_plasma_frequency(eq->get_ion_density(i, x, y, z), charge, mi, this->c, this->epsilon0);
                auto ic = build_cyclotron_frequency(charge, b_len, mi, this->c);
                auto denomi = 1.0 - ic*ic/w2;
                e11 = e11 - (wpi2/w2)/denomi;
                e12 = e12 + ((ic/w)*(wpi2/w2))/denomi;
                e33 = e33 + wpi2;
            }
            e12 = -1.0*e12;
            e33 = 1.0 - e33/w2;

            auto n = k_vec/w;
            auto b_hat = b_vec->unit();
            auto npara = b_hat->dot(n);
            auto npara2 = npara*npara;
            auto nperp = b_hat->cross(n)->length();
            auto nperp2 = nperp*nperp;

            elements e;
            e.w2 = w2;
            e.b_sq = b_vec->dot(b_vec);
            e.b_len = b_len;
            e.npara2 = npara2;
            e.nperp2 = nperp2;
            e.m11 = e11 - npara2;
            e.m12 = e12;
            e.m13 = npara*nperp;
            e.m22 = e11 - npara2 - nperp2;
            e.m33 = e33 - nperp2;
            return e;
        }
    public:
        virtual leaf_ptr D(leaf_ptr w, vector_ptr k_vec, leaf_ptr x, leaf_ptr y, leaf_ptr z, leaf_ptr,
                           equilibrium::shared<T, SAFE_MATH> &eq) {
            const elements e = build(w, k_vec, x, y, z, eq);
            return (e.m11*e.m22 - e.m12*e.m12)*e.m33 - e.m22*(e.m13*e.m13);
        }

///  What the reference's symbolic dD/d(axis) contains on top of the true derivative.
///
///  The reference's expression reducer rewrites ((a b)^2 c)/(d^2 b^4) as a^2 c/d^4 (it should be
///  a^2 c/(d^2 b^2); four plain variables reproduce it).  n_par^2 and
///  n_perp^2 are N^2/(B.B w^2); in the quotient rule of their derivative along a coordinate on which
///  B.B depends through a common factor s (the 1/R of the EFIT field components, coordinate z), the
///  term -n^2 d(B.B)/(B.B) comes out multiplied by w^2/((B.B)^2 s)  (s = R^8 for EFIT, given as 1/s).  The faulty form survives in the
///  copies of n_par^2 and n_perp^2 inside m11 and m22; m33 and m13 reduce along another path and are
///  right.  With K = -(d(B.B)/B.B) (w^2/((B.B)^2 s) - 1):
///      extra = K ((n_par^2 + n_perp^2) m13^2 - (n_par^2 m22 + m11 (n_par^2 + n_perp^2)) m33).
///  Checked against the reference's own kernels: dkz/dt agrees to 4e-15 (median) on the 64 + 16 states
///  of tests/golden/ref_rhs_cold_plasma_efit.npz and ref_defect_cold_plasma_efit.npz.
        virtual leaf_ptr reference_defect(leaf_ptr w, vector_ptr k_vec, leaf_ptr x, leaf_ptr y, leaf_ptr z, leaf_ptr,
                                          equilibrium::shared<T, SAFE_MATH> &eq, int &axis) {
            const auto defect = eq->get_reducer_defect(x, y, z);
            axis = defect.axis;
            if (axis < 0) return leaf_ptr();
            const elements e = build(w, k_vec, x, y, z, eq);
            leaf_ptr coordinate = axis == 0 ? x : (axis == 1 ? y : z);
            auto db_sq = defect.b_sq_derivative.get() ? defect.b_sq_derivative : e.b_sq->df(coordinate);
//  1/(B.B) and 1/(B.B)^2 through 1/|B|, which the kernel has as the reciprocal root of B.B.
            auto inv_b_sq = (1.0/e.b_len)/e.b_len;
            auto K = -1.0*(db_sq*inv_b_sq)*(e.w2*(inv_b_sq*inv_b_sq)*defect.inverse_scale - 1.0);
            auto n2 = e.npara2 + e.nperp2;
            return K*(n2*(e.m13*e.m13) - (e.npara2*e.m22 + e.m11*n2)*e.m33);
        }
    };

//------------------------------------------------------------------------------
//  Absorption path (dispersion.hpp:245-305, 1010-1098, 1209-1290 of the reference).
//
//  The reference runs this stage in std::complex<T>.  Ray state is real, so every quantity is real
//  except the plasma dispersion function Z(zeta); a complex value is therefore carried as a pair of
//  real graphs and the few complex operations (Z, 1/Z, scaling by a real) are written out.  The real
//  arithmetic below is what the complex arithmetic of the reference reduces to for zero imaginary
//  parts, up to rounding in complex division.
//------------------------------------------------------------------------------
    struct complex_leaf {
        leaf_ptr re, im;
    };

///  Z(zeta) = -sqrt(pi) exp(-zeta^2) (erfi(zeta) - i)  (z_erfi, dispersion.hpp:289-305).
    template<typename T=double, bool SAFE_MATH=false>
    class z_erfi {
    public:
        complex_leaf Z(leaf_ptr zeta) {
            auto scale = -std::sqrt(M_PI)*graph::exp(-1.0*(zeta*zeta));
            return {scale*graph::erfi(zeta), -1.0*scale};
        }
    };

///  Cold plasma dispersion function in the form the hot plasma expansion reduces to for
///  v_th -> 0 (dispersion.hpp:1010-1098); real, usable as an ordinary dispersion function.
    template<typename T=double, bool SAFE_MATH=false>
    class cold_plasma_expansion : public physics<T, SAFE_MATH> {
    public:
        virtual leaf_ptr D(leaf_ptr w, vector_ptr k_vec, leaf_ptr x, leaf_ptr y, leaf_ptr z, leaf_ptr,
                           equilibrium::shared<T, SAFE_MATH> &eq) {
            auto b_vec = eq->get_magnetic_field(x, y, z);
            auto b_len = b_vec->length();
            auto b_hat = b_vec/b_len;
            auto ec = build_cyclotron_frequency(this->q, b_len, this->me, this->c);
            auto wpe2 = build_plasma_frequency(eq->get_electron_density(x, y, z), this->q, this->me, this->c, this->epsilon0);

            auto P = wpe2/(w*w);
            auto q = P/(2.0*(1.0 + ec/w));

            auto n = k_vec/w;
            auto n2 = n->dot(n);
            auto npara = n->dot(b_hat);
            auto npara2 = npara*npara;
            auto nperp = b_hat->cross(n);
            auto nperp2 = nperp->dot(nperp);

            auto q_func = 1.0 - 2.0*q;
            auto n_func = n2 + npara2;
            auto p_func = 1.0 - P;

            auto gamma1 = (1.0 - q)*(n2*nperp2) + p_func*(n2*npara2 - (1.0 - q)*n_func) + q_func*(p_func - nperp2);
            auto gamma0 = nperp2*(n2 - 2.0*q_func) + p_func*(2.0*q_func - n_func);
            return -1.0*P/2.0*(1.0 + ec/w)*gamma0 + (1.0 - ec*ec/(w*w))*gamma1;
        }
    };

///  Weakly relativistic / warm correction near the electron cyclotron fundamental
///  (hot_plasma_expansion, dispersion.hpp:1209-1290).  Complex through Z only.
    template<typename T=double, class Z=z_erfi<T, false>, bool SAFE_MATH=false>
    class hot_plasma_expansion : public physics<T, SAFE_MATH> {
    private:
        Z z_function;
    public:
///  Not a ray-tracing Hamiltonian: the real part alone is returned through the common interface.
        virtual leaf_ptr D(leaf_ptr w, vector_ptr k_vec, leaf_ptr x, leaf_ptr y, leaf_ptr z, leaf_ptr t,
                           equilibrium::shared<T, SAFE_MATH> &eq) {
            return D_complex(w, k_vec, x, y, z, t, eq).re;
        }
        complex_leaf D_complex(leaf_ptr w, vector_ptr k_vec, leaf_ptr x, leaf_ptr y, leaf_ptr z, leaf_ptr,
                               equilibrium::shared<T, SAFE_MATH> &eq) {
            auto b_vec = eq->get_magnetic_field(x, y, z);
            auto b_hat = b_vec->unit();
            auto b_len = b_vec->length();
            auto te = eq->get_electron_temperature(x, y, z);
            auto ve = graph::sqrt(static_cast<T> (2.0)*this->q*te/this->me);
            auto ec = build_cyclotron_frequency(this->q, b_len, this->me, this->c);
            auto wpe2 = build_plasma_frequency(eq->get_electron_density(x, y, z), this->q, this->me, this->c, this->epsilon0);

            auto P = wpe2/(w*w);
            auto q = P/(2.0*(1.0 + ec/w));

            auto n = k_vec/w;
            auto n2 = n->dot(n);
            auto npara = b_hat->dot(n);
            auto npara2 = npara*npara;
            auto nperp = b_hat->cross(n);
            auto nperp2 = nperp->dot(nperp);
            auto vtnorm = ve/this->c;

            auto zeta = (1.0 - ec/w)/(npara*vtnorm);
            auto Zf = z_function.Z(zeta);

            auto q_func = 1.0 - 2.0*q;
            auto n_func = n2 + npara2;
            auto p_func = 1.0 - P;

            auto gamma5 = P*(n2*npara2 - (1.0 - q)*n_func + q_func);
            auto gamma2 = P*w/ec*nperp2*(n2 - q_func) + P*P*w*w/(4.0*ec*ec)*(n_func - 2.0*q_func)*nperp2/npara2;
            auto gamma1 = (1.0 - q)*(n2*nperp2) + p_func*(n2*npara2 - (1.0 - q)*n_func) + q_func*(p_func - nperp2);

            auto amplitude = -1.0*(1.0 + ec/w)*npara*vtnorm*
                             (gamma1 + gamma2 + nperp2/(2.0*npara)*(w*w/(ec*ec))*vtnorm*zeta*gamma5);
//  1/Z = conj(Z)/|Z|^2.  Far from the resonance (|zeta| > 27.2) exp(-zeta^2) underflows, Z = 0 and
//  1/Z is 0/0.  The reference leaves that to its SAFE_MATH rules (products with an exact zero are
//  zero, NaN is stored as 0: cpu_context.hpp:533-544) and to the fast-math of its JIT compiler; what
//  its kernels deliver there is D = 0, i.e. no damping, which is also the analytic limit
//  (Z -> -1/zeta).  Stated explicitly here.
            auto norm = Zf.re*Zf.re + Zf.im*Zf.im;
            return {graph::nan_to_zero(amplitude*(Zf.re/norm + zeta)),
                    graph::nan_to_zero(amplitude*(-1.0*Zf.im/norm))};
        }
    };

///  See dispersion_interface: whether the k_vec correction term is applied (default: no, like
///  the reference's effective behaviour).
    inline bool &kvec_correction() {
        static thread_local bool on = false;
        return on;
    }

///  Whether the ray equations reproduce the reference where its symbolic derivative is defective
///  (cold_plasma in the EFIT field: dD/dz, see cold_plasma::reference_defect).  Default: yes -- results
///  identical to the reference's; false gives the true derivative.  GFB_TRUE_DERIVATIVES=1 in the
///  environment or the tracer option reference_defects=0 switch it off.
    inline bool &reference_defects() {
        static thread_local bool on = std::getenv("GFB_TRUE_DERIVATIVES") == nullptr;
        return on;
    }

///  Whether dispersion_interface differentiates D with one reverse sweep (graph::gradient)
///  instead of seven forward df() calls.
    inline bool &reverse_mode() {
        static thread_local bool on = std::getenv("GFB_FORWARD_MODE") == nullptr;
        return on;
    }

    template<class D>
    concept function = std::is_base_of<dispersion_function<typename D::base, D::safe_math>, D>::value;

//------------------------------------------------------------------------------
///  Ray equations from a dispersion function (dispersion.hpp:1369-1434).
//------------------------------------------------------------------------------
    template<function DISPERSION_FUNCTION>
    class dispersion_interface {
    protected:
        typedef typename DISPERSION_FUNCTION::base T;
        static constexpr bool SAFE_MATH = DISPERSION_FUNCTION::safe_math;
        vector_ptr k_vec;
        leaf_ptr D;
        leaf_ptr dxdt, dydt, dzdt, dkxdt, dkydt, dkzdt, dsdt;
    public:
        dispersion_interface(leaf_ptr w, leaf_ptr kx, leaf_ptr ky, leaf_ptr kz,
                             leaf_ptr x, leaf_ptr y, leaf_ptr z, leaf_ptr t,
                             equilibrium::shared<T, SAFE_MATH> &eq) :
        k_vec(kx*eq->get_esup1(x, y, z) + ky*eq->get_esup2(x, y, z) + kz*eq->get_esup3(x, y, z)),
        D(DISPERSION_FUNCTION().D(w, k_vec, x, y, z, t, eq)) {
//  Correction for k_vec depending on the coordinates (curvilinear equilibria).
            auto dkdx = k_vec->df(x);
            auto dkdy = k_vec->df(y);
            auto dkdz = k_vec->df(z);
            auto dDdk_vec = graph::vector(D->df(k_vec->get_x()), D->df(k_vec->get_y()), D->df(k_vec->get_z()));

            leaf_ptr dDdw, dDdkx, dDdky, dDdkz, dDdx, dDdy, dDdz;
            if (reverse_mode()) {
//  One backward sweep for all seven derivatives (graph::gradient); same values as df().
                auto g = graph::gradient(D, {w, kx, ky, kz, x, y, z});
                dDdw = g[0]; dDdkx = g[1]; dDdky = g[2]; dDdkz = g[3]; dDdx = g[4]; dDdy = g[5]; dDdz = g[6];
            } else {
                dDdw = D->df(w);
                dDdkx = D->df(kx);
                dDdky = D->df(ky);
                dDdkz = D->df(kz);
                dDdx = D->df(x);
                dDdy = D->df(y);
                dDdz = D->df(z);
            }

            if (reference_defects()) {
                int axis = -1;
                auto extra = DISPERSION_FUNCTION().reference_defect(w, k_vec, x, y, z, t, eq, axis);
                if (extra.get()) {
                    if (axis == 0) dDdx = dDdx + extra;
                    if (axis == 1) dDdy = dDdy + extra;
                    if (axis == 2) dDdz = dDdz + extra;
                }
            }

            if (graph::pseudo_variable_cast(x).get()) {
                dkdx = dkdx->remove_pseudo();
                dkdy = dkdy->remove_pseudo();
                dkdz = dkdz->remove_pseudo();
                dDdk_vec = dDdk_vec->remove_pseudo();
                dDdw = dDdw->remove_pseudo();
                dDdkx = dDdkx->remove_pseudo();
                dDdky = dDdky->remove_pseudo();
                dDdkz = dDdkz->remove_pseudo();
                dDdx = dDdx->remove_pseudo();
                dDdy = dDdy->remove_pseudo();
                dDdz = dDdz->remove_pseudo();
            }

            dxdt = -dDdkx/dDdw;
            dydt = -dDdky/dDdw;
            dzdt = -dDdkz/dDdw;
//  The reference subtracts dD/dk_vec . dk_vec/dx so that the coordinate derivative is taken at
//  fixed Cartesian k_vec (dispersion.hpp:1427-1429).  In the reference that term is identically
//  zero in practice: for Cartesian equilibria dk_vec/dx = 0, and for curvilinear ones (VMEC) its
//  reducer rewrites n = k_vec/w so that the k_vec component nodes no longer occur inside D and
//  D->df(k_vec->get_x()) reduces to zero.  Measured: with the term dropped this back end matches
//  the reference's VMEC right-hand side to 1e-15 on every component, with it kept dk/dt differs by
//  O(1) (tests/golden/ref_rhs_ordinary_wave_vmec.npz).  Reproduced, not fixed;
//  set dispersion::kvec_correction() = true for the formula as written.
            if (kvec_correction()) {
                dkxdt = (dDdx - dDdk_vec->dot(dkdx))/dDdw;
                dkydt = (dDdy - dDdk_vec->dot(dkdy))/dDdw;
                dkzdt = (dDdz - dDdk_vec->dot(dkdz))/dDdw;
            } else {
                dkxdt = dDdx/dDdw;
                dkydt = dDdy/dDdw;
                dkzdt = dDdz/dDdw;
            }
            dsdt = graph::vector(dxdt, dydt, dzdt)->length();
        }

///  Newton solve of D = 0 for one of the inputs (dispersion.hpp:1452-1475).
        leaf_ptr solve(leaf_ptr x, graph::input_nodes<T, SAFE_MATH> inputs, const size_t index=0,
                       const T tolerance = 1.0E-30, const size_t max_iterations = 1000,
                       const solver::newton_mode mode = solver::newton_mode::ensemble) {
            auto x_var = graph::variable_cast(x);
            workflow::manager<T, SAFE_MATH> work(index);
            solver::newton<T, SAFE_MATH> (work, {x}, inputs, D, graph::shared_random_state<T, SAFE_MATH> (),
                                          tolerance, max_iterations, 1.0, mode);
            work.compile();
            work.run();
            work.copy_to_host(x, x_var->data());
            return D*D;
        }

        leaf_ptr get_residual() { return D*D; }
        leaf_ptr get_d() { return D; }
        leaf_ptr get_dsdt() { return dsdt; }
        leaf_ptr get_dxdt() { return dxdt; }
        leaf_ptr get_dydt() { return dydt; }
        leaf_ptr get_dzdt() { return dzdt; }
        leaf_ptr get_dkxdt() { return dkxdt; }
        leaf_ptr get_dkydt() { return dkydt; }
        leaf_ptr get_dkzdt() { return dkzdt; }
        void print_dispersion() { D->to_latex(); std::cout << std::endl; }
    };
}

#endif /* gfb_graph_dispersion_hpp */
