//------------------------------------------------------------------------------
//  equilibrium.hpp -- plasma equilibria as expression graphs.
//
//  Mirrors the interface of /root/reference/graph_framework/equilibrium.hpp:
//  equilibrium::generic (:214-470) and the concrete equilibria
//  no_magnetic_field (:482), slab (:611), slab_density (:735), slab_field (:864),
//  gaussian_density (:991), efit (:1146-1616, make_efit :1628-1854) and
//  vmec (:1868-2413, make_vmec :2424-2640).
//
//  Table files: the reference reads netCDF-4; this back end reads the GFBT
//  container (graph_framework_b200/tools/gfbt.py) that the converter writes from
//  the same files, variable names unchanged.
//
//  Reference quirks that change numbers are reproduced on purpose (SURVEY.md
//  Appendix B): the density spline's first two tables are the temperature's
//  (equilibrium.hpp:1478), ion density = electron temperature (:1361), the
//  efit-local charge 1.60218E-19 (:1359).
//------------------------------------------------------------------------------
#ifndef gfb_graph_equilibrium_hpp
#define gfb_graph_equilibrium_hpp

#include <cstdio>

#include "newton.hpp"
#include "vector.hpp"

namespace equilibrium {
    using graph::leaf_ptr;
    using graph::vector_ptr;

//------------------------------------------------------------------------------
///  GFBT reader.
//------------------------------------------------------------------------------
    struct table_file {
        std::map<std::string, std::vector<double>> vars;
        std::map<std::string, std::vector<size_t>> dims;

        explicit table_file(const std::string &path) {
            FILE *f = std::fopen(path.c_str(), "rb");
            if (!f) {
                std::cerr << "Cannot open table file " << path << std::endl;
                std::exit(-1);
            }
            char magic[6];
            uint32_t n = 0;
            bool ok = std::fread(magic, 1, 6, f) == 6 && !std::memcmp(magic, "GFBT1\n", 6) && std::fread(&n, 4, 1, f) == 1;
            for (uint32_t i = 0; ok && i < n; i++) {
                uint32_t len = 0, rank = 0;
                ok = std::fread(&len, 4, 1, f) == 1;
                std::string name(len, '\0');
                ok = ok && std::fread(name.data(), 1, len, f) == len && std::fread(&rank, 4, 1, f) == 1;
                std::vector<size_t> d;
                size_t count = 1;
                for (uint32_t r = 0; ok && r < rank; r++) {
                    uint64_t e = 0;
                    ok = std::fread(&e, 8, 1, f) == 1;
                    d.push_back(e);
                    count *= e;
                }
                std::vector<double> v(count);
                ok = ok && std::fread(v.data(), 8, count, f) == count;
                vars[name] = std::move(v);
                dims[name] = std::move(d);
            }
            std::fclose(f);
            if (!ok) {
                std::cerr << "Malformed table file " << path << " (expected GFBT1; convert netCDF files with "
                             "graph_framework_b200.tools.gfbt.nc_to_gfbt)." << std::endl;
                std::exit(-1);
            }
        }
        const std::vector<double> &get(const std::string &name) const {
            auto it = vars.find(name);
            if (it == vars.end()) {
                std::cerr << "Table file has no variable " << name << std::endl;
                std::exit(-1);
            }
            return it->second;
        }
        double scalar(const std::string &name) const { return get(name).at(0); }
        size_t dim(const std::string &name) const { return static_cast<size_t> (get("dim:" + name).at(0)); }
    };

//------------------------------------------------------------------------------
///  Interface (equilibrium.hpp:214-470).
//------------------------------------------------------------------------------
///  The grid of the largest coefficient tables of an equilibrium, for keeping rays sorted by cell
///  on the device (gfb_bin_rays / gfb_bin_rays_rz in include/gfb200.h).  dims = 0: no tables.
    struct cell_grid {
        int dims = 0;                   ///< 1: uniform in the first coordinate; 2: (R, Z) with R = sqrt(x^2 + y^2)
        double lo[2] = {0.0, 0.0}, hi[2] = {1.0, 1.0};
        unsigned cells[2] = {0, 0};
        double cell_length = 0.0;       ///< length a ray travels across one cell (a ray moves at most dt per step, c = 1)
///  A re-sort costs about as much as `min_steps`/25 steps of the kernels that use these tables, so
///  checking the order more often than every min_steps steps cannot pay.
        size_t min_steps = 20;
///  Steps after which the order is worth checking: about half a cell of travel, within [min_steps, 5000].
        size_t drift_steps(const double dt) const {
            if (!(cell_length > 0.0) || !(dt > 0.0)) return 1000;
            return static_cast<size_t> (std::min(5000.0, std::max(static_cast<double> (min_steps), 0.5*cell_length/dt)));
        }
    };

    template<typename T=double, bool SAFE_MATH=false>
    class generic {
    protected:
        const std::vector<T> ion_masses;
        const std::vector<uint8_t> ion_charges;
    public:
        generic(const std::vector<T> &masses, const std::vector<uint8_t> &charges) :
        ion_masses(masses), ion_charges(charges) {
            assert(ion_masses.size() == ion_charges.size() && "Masses and charges need the same number of elements.");
        }
        virtual ~generic() {}
        size_t get_num_ion_species() const { return ion_masses.size(); }
        T get_ion_mass(const size_t index) const { return ion_masses.at(index); }
        uint8_t get_ion_charge(const size_t index) const { return ion_charges.at(index); }

        virtual leaf_ptr get_electron_density(leaf_ptr x, leaf_ptr y, leaf_ptr z) = 0;
        virtual leaf_ptr get_ion_density(const size_t index, leaf_ptr x, leaf_ptr y, leaf_ptr z) = 0;
        virtual leaf_ptr get_electron_temperature(leaf_ptr x, leaf_ptr y, leaf_ptr z) = 0;
        virtual leaf_ptr get_ion_temperature(const size_t index, leaf_ptr x, leaf_ptr y, leaf_ptr z) = 0;
        virtual vector_ptr get_magnetic_field(leaf_ptr x, leaf_ptr y, leaf_ptr z) = 0;
        virtual leaf_ptr get_characteristic_field(const size_t device_number=0) = 0;

        virtual vector_ptr get_esup1(leaf_ptr, leaf_ptr, leaf_ptr) { return graph::vector(1.0, 0.0, 0.0); }
        virtual vector_ptr get_esup2(leaf_ptr, leaf_ptr, leaf_ptr) { return graph::vector(0.0, 1.0, 0.0); }
        virtual vector_ptr get_esup3(leaf_ptr, leaf_ptr, leaf_ptr) { return graph::vector(0.0, 0.0, 1.0); }
        virtual leaf_ptr get_x(leaf_ptr x1, leaf_ptr, leaf_ptr) { return x1; }
        virtual leaf_ptr get_y(leaf_ptr, leaf_ptr x2, leaf_ptr) { return x2; }
        virtual leaf_ptr get_z(leaf_ptr, leaf_ptr, leaf_ptr x3) { return x3; }
///  Extension (no reference counterpart): see cell_grid.
        virtual cell_grid get_cell_grid() const { return cell_grid(); }
///  Reference compatibility (dispersion::cold_plasma::reference_defect): the coordinate along which
///  the reference's reducer mis-cancels the derivative of B.B in this field, and the factor involved.
        struct reducer_defect {
            int axis = -1;              ///< 0, 1, 2 = x, y, z; -1: none known
            leaf_ptr inverse_scale;     ///< 1/s, see dispersion::cold_plasma::reference_defect
            leaf_ptr b_sq_derivative;   ///< d(B.B)/d(axis) where the equilibrium has a cheaper form than df(); may be null
        };
        virtual reducer_defect get_reducer_defect(leaf_ptr, leaf_ptr, leaf_ptr) { return reducer_defect(); }
    };

    template<typename T=double, bool SAFE_MATH=false>
    using shared = std::shared_ptr<generic<T, SAFE_MATH>>;

///  Deuterium ion, charge one: the species every reference equilibrium uses.
    constexpr double deuterium_mass = 3.34449469E-27;

//------------------------------------------------------------------------------
///  Analytic equilibria share one implementation parameterised by their
///  density, temperature and field profiles in x.
//------------------------------------------------------------------------------
    template<typename T=double, bool SAFE_MATH=false>
    class analytic : public generic<T, SAFE_MATH> {
    public:
        using profile = std::function<leaf_ptr(leaf_ptr, leaf_ptr, leaf_ptr)>;
        using field = std::function<vector_ptr(leaf_ptr, leaf_ptr, leaf_ptr)>;
    private:
        profile density, temperature;
        field b;
    public:
        analytic(profile density, profile temperature, field b) :
        generic<T, SAFE_MATH> ({deuterium_mass}, {1}), density(density), temperature(temperature), b(b) {}
        virtual leaf_ptr get_electron_density(leaf_ptr x, leaf_ptr y, leaf_ptr z) final { return density(x, y, z); }
        virtual leaf_ptr get_ion_density(const size_t, leaf_ptr x, leaf_ptr y, leaf_ptr z) final { return density(x, y, z); }
        virtual leaf_ptr get_electron_temperature(leaf_ptr x, leaf_ptr y, leaf_ptr z) final { return temperature(x, y, z); }
        virtual leaf_ptr get_ion_temperature(const size_t, leaf_ptr x, leaf_ptr y, leaf_ptr z) final { return temperature(x, y, z); }
        virtual vector_ptr get_magnetic_field(leaf_ptr x, leaf_ptr y, leaf_ptr z) final { return b(x, y, z); }
        virtual leaf_ptr get_characteristic_field(const size_t=0) final { return graph::one(); }
    };

    namespace profiles {
        inline leaf_ptr linear(leaf_ptr x, const double amplitude, const double slope) {
            return graph::constant(amplitude)*(graph::constant(slope)*x + graph::one());
        }
    }

///  ne = 1e19 (0.1 x + 1), Te = 1000, B = 0  (equilibrium.hpp:482-600).
    template<typename T=double, bool SAFE_MATH=false>
    shared<T, SAFE_MATH> make_no_magnetic_field() {
        return std::make_shared<analytic<T, SAFE_MATH>> (
            [] (leaf_ptr x, leaf_ptr, leaf_ptr) { return profiles::linear(x, 1.0E19, 0.1); },
            [] (leaf_ptr, leaf_ptr, leaf_ptr) { return graph::constant(1000.0); },
            [] (leaf_ptr, leaf_ptr, leaf_ptr) { return graph::vector(0.0, 0.0, 0.0); });
    }
///  ne = 1e19, Te = 1000, B = (0, 0, 0.1 x + 1)  (equilibrium.hpp:611-724).
    template<typename T=double, bool SAFE_MATH=false>
    shared<T, SAFE_MATH> make_slab() {
        return std::make_shared<analytic<T, SAFE_MATH>> (
            [] (leaf_ptr, leaf_ptr, leaf_ptr) { return graph::constant(1.0E19); },
            [] (leaf_ptr, leaf_ptr, leaf_ptr) { return graph::constant(1000.0); },
            [] (leaf_ptr x, leaf_ptr, leaf_ptr) { return graph::vector(0.0, 0.0, 0.1*x + 1.0); });
    }
///  ne = 1e19 (0.1 x + 1), Te = 1000, B = (0, 0, 1)  (equilibrium.hpp:735-853).
    template<typename T=double, bool SAFE_MATH=false>
    shared<T, SAFE_MATH> make_slab_density() {
        return std::make_shared<analytic<T, SAFE_MATH>> (
            [] (leaf_ptr x, leaf_ptr, leaf_ptr) { return profiles::linear(x, 1.0E19, 0.1); },
            [] (leaf_ptr, leaf_ptr, leaf_ptr) { return graph::constant(1000.0); },
            [] (leaf_ptr, leaf_ptr, leaf_ptr) { return graph::vector(0.0, 0.0, 1.0); });
    }
///  ne = 1e19 (0.01 x + 1), Te = 2000 (0.01 x + 1), B = (0, 0, 0.01 x + 1)  (equilibrium.hpp:864-980).
    template<typename T=double, bool SAFE_MATH=false>
    shared<T, SAFE_MATH> make_slab_field() {
        return std::make_shared<analytic<T, SAFE_MATH>> (
            [] (leaf_ptr x, leaf_ptr, leaf_ptr) { return profiles::linear(x, 1.0E19, 0.01); },
            [] (leaf_ptr x, leaf_ptr, leaf_ptr) { return profiles::linear(x, 2000.0, 0.01); },
            [] (leaf_ptr x, leaf_ptr, leaf_ptr) { return graph::vector(0.0, 0.0, 0.01*x + 1.0); });
    }
///  ne = 1e19 exp((x^2 + y^2)/-0.2), Te = 1000, B = (1, 0, 0)  (equilibrium.hpp:991-1108).
    template<typename T=double, bool SAFE_MATH=false>
    shared<T, SAFE_MATH> make_gaussian_density() {
        return std::make_shared<analytic<T, SAFE_MATH>> (
            [] (leaf_ptr x, leaf_ptr y, leaf_ptr) {
                return graph::constant(1.0E19)*graph::exp((x*x + y*y)/graph::constant(-0.2));
            },
            [] (leaf_ptr, leaf_ptr, leaf_ptr) { return graph::constant(1000.0); },
            [] (leaf_ptr, leaf_ptr, leaf_ptr) { return graph::vector(1.0, 0.0, 0.0); });
    }

//------------------------------------------------------------------------------
///  Cubic spline in the *un-normalised* argument: the cell's coefficients are
///  rewritten on the host so that  p(x) = ((c3 x + c2) x + c1) x + c0  with x the
///  physical coordinate (equilibrium.hpp:1121-1133).  The rewrite happens in
///  table space (fold_tables_scope), giving four tables on the same cells.
//------------------------------------------------------------------------------
    inline std::array<leaf_ptr, 4> fold_1D_spline(graph::output_nodes<> c, const double scale, const double offset) {
        graph::fold_tables_scope fold;
        const double s2 = scale*scale, s3 = scale*scale*scale;
        std::array<leaf_ptr, 4> k;
        k[3] = c[3]/s3;
        k[2] = c[2]/s2 - 3.0*offset*c[3]/s3;
        k[1] = c[1]/scale - 2.0*offset*c[2]/s2 + 3.0*offset*offset*c[3]/s3;
        k[0] = c[0] - offset*c[1]/scale + offset*offset*c[2]/s2 - offset*offset*offset*c[3]/s3;
        return k;
    }
    inline leaf_ptr build_1D_spline(graph::output_nodes<> c, leaf_ptr x, const double scale, const double offset) {
        const auto k = fold_1D_spline(c, scale, offset);
        return graph::fma(graph::fma(graph::fma(k[3], x, k[2]), x, k[1]), x, k[0]);
    }

///  Whether tabulated equilibria build their cubics as graph::spline_1d / spline_2d nodes (one node per
///  evaluated quantity, closed under df(); all derivative orders a kernel needs come out of one pass over
///  the cell's coefficients) or as the reference's Horner chains of piecewise nodes (build_1D_spline,
///  build_psi), differentiated link by link.  Same mathematics; GFB_SPLINE_NODES=0 selects the chains.
    inline bool &spline_nodes() {
        static thread_local bool on = !(std::getenv("GFB_SPLINE_NODES") && std::getenv("GFB_SPLINE_NODES")[0] == '0');
        return on;
    }

//------------------------------------------------------------------------------
///  EFIT tokamak equilibrium: bicubic psi(R, Z), cubic profiles in psi.
//------------------------------------------------------------------------------
    template<typename T=double, bool SAFE_MATH=false>
    class efit final : public generic<T, SAFE_MATH> {
    public:
        struct tables {
            double psimin, dpsi, rmin, dr, zmin, dz;
            double te_scale, ne_scale, pres_scale;
            size_t num_cols;
            std::array<std::vector<double>, 4> te, ne, pres, fpol;
            std::array<std::array<std::vector<double>, 4>, 4> psi;     // psi[i][j] = c_ij, i: power of R, j: power of Z
        };
    private:
        const tables tab;
        leaf_ptr x_cache, y_cache, z_cache;
        leaf_ptr ne_cache, ni_cache, te_cache, ti_cache, psi_cache, r_cache, fpol_cache;
        vector_ptr b_cache;

        leaf_ptr profile(const std::array<std::vector<double>, 4> &c, leaf_ptr psi) {
            graph::output_nodes<> raw = {graph::piecewise_1D(c[0], psi, tab.dpsi, tab.psimin),
                                         graph::piecewise_1D(c[1], psi, tab.dpsi, tab.psimin),
                                         graph::piecewise_1D(c[2], psi, tab.dpsi, tab.psimin),
                                         graph::piecewise_1D(c[3], psi, tab.dpsi, tab.psimin)};
            if (spline_nodes()) return graph::spline_1d(fold_1D_spline(raw, tab.dpsi, tab.psimin), psi);
            return build_1D_spline(raw, psi, tab.dpsi, tab.psimin);
        }

///  equilibrium.hpp:1279-1313.
        leaf_ptr build_psi(leaf_ptr r, leaf_ptr z) {
            std::array<leaf_ptr, 4> c;
            std::array<std::array<leaf_ptr, 4>, 4> folded;
            for (size_t i = 0; i < 4; i++) {
                graph::output_nodes<> row;
                for (size_t j = 0; j < 4; j++) {
                    row.push_back(graph::piecewise_2D(tab.psi[i][j], tab.num_cols, r, tab.dr, tab.rmin, z, tab.dz, tab.zmin));
                }
                if (spline_nodes()) folded[i] = fold_1D_spline(row, tab.dz, tab.zmin);
                else c[i] = build_1D_spline(row, z, tab.dz, tab.zmin);
            }
            if (spline_nodes()) return graph::spline_2d(folded, r, z);
            auto r_norm = (r - tab.rmin)/tab.dr;
            return ((c[3]*r_norm + c[2])*r_norm + c[1])*r_norm + c[0];
        }

///  equilibrium.hpp:1324-1384.
        void set_cache(leaf_ptr x, leaf_ptr y, leaf_ptr z) {
            if (x->is_match(x_cache) && y->is_match(y_cache) && z->is_match(z_cache)) return;
            x_cache = x;
            y_cache = y;
            z_cache = z;
            auto r = graph::sqrt(x*x + y*y);
            r_cache = r;
            psi_cache = build_psi(r, z);
            ne_cache = graph::constant(tab.ne_scale)*profile(tab.ne, psi_cache);
            te_cache = graph::constant(tab.te_scale)*profile(tab.te, psi_cache);
            auto pressure = graph::constant(tab.pres_scale)*profile(tab.pres, psi_cache);
            auto q = graph::constant(1.60218E-19);
            ni_cache = te_cache;
            ti_cache = (pressure - ne_cache*te_cache*q)/(ni_cache*q);

            auto phi = graph::atan(x, y);
            auto br = psi_cache->df(z)/r;
            fpol_cache = profile(tab.fpol, psi_cache);
            auto bp = fpol_cache/r;
            auto bz = -psi_cache->df(r)/r;
            auto cos = graph::cos(phi);
            auto sin = graph::sin(phi);
            b_cache = graph::vector(br*cos - bp*sin, br*sin + bp*cos, bz);
        }

    public:
        efit(const tables &t) : generic<T, SAFE_MATH> ({deuterium_mass}, {1}), tab(t) {
            x_cache = y_cache = z_cache = graph::zero();
        }
        virtual leaf_ptr get_electron_density(leaf_ptr x, leaf_ptr y, leaf_ptr z) { set_cache(x, y, z); return ne_cache; }
        virtual leaf_ptr get_ion_density(const size_t, leaf_ptr x, leaf_ptr y, leaf_ptr z) { set_cache(x, y, z); return ni_cache; }
        virtual leaf_ptr get_electron_temperature(leaf_ptr x, leaf_ptr y, leaf_ptr z) { set_cache(x, y, z); return te_cache; }
        virtual leaf_ptr get_ion_temperature(const size_t, leaf_ptr x, leaf_ptr y, leaf_ptr z) { set_cache(x, y, z); return ti_cache; }
        virtual vector_ptr get_magnetic_field(leaf_ptr x, leaf_ptr y, leaf_ptr z) { set_cache(x, y, z); return b_cache; }
        leaf_ptr get_psi(leaf_ptr x, leaf_ptr y, leaf_ptr z) { set_cache(x, y, z); return psi_cache; }
///  Every field component carries 1/R (equilibrium.hpp:1363-1381): B.B = G/R^2, and the reducer's
///  faulty cancellation along z leaves R^8 behind (measured: exact to 10 digits on every state).
        virtual typename generic<T, SAFE_MATH>::reducer_defect get_reducer_defect(leaf_ptr x, leaf_ptr y, leaf_ptr z) {
            typename generic<T, SAFE_MATH>::reducer_defect d;
            set_cache(x, y, z);
//  1/R^8 from the 1/R the kernel already has (every division by the root node R is a multiplication by its rsqrt).
            auto inv_r2 = (1.0/r_cache)/r_cache;
            auto inv_r4 = inv_r2*inv_r2;
            d.axis = 2;
            d.inverse_scale = inv_r4*inv_r4;
//  B.B = (psi_Z^2 + F^2 + psi_R^2)/R^2 in cylindrical components (the rotation by phi drops out), so
//  d(B.B)/dz = 2 (psi_Z psi_ZZ + F F' psi_Z + psi_R psi_RZ)/R^2: every factor is already in the kernel.
            auto psi_z = psi_cache->df(z);
            auto psi_r = psi_cache->df(r_cache);
            auto f = fpol_cache;
//  (two divisions by R: the kernel has 1/R from the square root already, 1/R^2 would be a new reciprocal)
            d.b_sq_derivative = 2.0*(psi_z*psi_z->df(z) + f*f->df(psi_cache)*psi_z + psi_r*psi_r->df(z))/r_cache/r_cache;
            return d;
        }
///  The psi(R, Z) tables: numr x numz cells.  A ray at the speed of light crosses one cell in
///  dr/dt steps; the hint assumes the reference's example step (2e-5).
        virtual cell_grid get_cell_grid() const {
            cell_grid g;
            const size_t num_rows = tab.psi[0][0].size()/tab.num_cols;
            g.dims = 2;
            g.lo[0] = tab.rmin; g.hi[0] = tab.rmin + tab.dr*static_cast<double> (num_rows); g.cells[0] = static_cast<unsigned> (num_rows);
            g.lo[1] = tab.zmin; g.hi[1] = tab.zmin + tab.dz*static_cast<double> (tab.num_cols); g.cells[1] = static_cast<unsigned> (tab.num_cols);
            g.cell_length = std::min(tab.dr, tab.dz);
            g.min_steps = 200;                  // EFIT ray steps are cheap (~0.08 us per ray): a sort is worth ~4 of them
            return g;
        }

///  |B| on the magnetic axis, found by a damped Newton search (step 0.1) for the
///  minimum of the normalised flux starting from (1.7, 0, 0)  (equilibrium.hpp:1584-1615).
        virtual leaf_ptr get_characteristic_field(const size_t device_number=0) final {
            auto x_axis = graph::variable(1, "x");
            auto y_axis = graph::variable(1, "y");
            auto z_axis = graph::variable(1, "z");
            x_axis->set(1.7);
            y_axis->set(0.0);
            z_axis->set(0.0);
            auto b_mod = get_magnetic_field(x_axis, y_axis, z_axis)->length();
            graph::input_nodes<> inputs {x_axis, y_axis, z_axis};
            workflow::manager<T, SAFE_MATH> work(device_number);
            solver::newton<T, SAFE_MATH> (work, {x_axis, z_axis}, inputs, (psi_cache - tab.psimin)/tab.dpsi,
                                          graph::shared_random_state<T, SAFE_MATH> (), 1.0E-30, 1000, 0.1);
            work.add_item(inputs, {b_mod}, {}, graph::shared_random_state<T, SAFE_MATH> (), "bmod_at_axis", 1);
            work.compile();
            work.run();
            T result;
            work.copy_to_host(b_mod, &result);
            return graph::constant(result);
        }
    };

///  Load EFIT tables (variable names of equilibrium.hpp:1628-1854) from a GFBT file.
    inline efit<>::tables load_efit_tables(const std::string &spline_file) {
        table_file f(spline_file);
        efit<>::tables t;
        t.rmin = f.scalar("rmin"); t.dr = f.scalar("dr");
        t.zmin = f.scalar("zmin"); t.dz = f.scalar("dz");
        t.psimin = f.scalar("psimin"); t.dpsi = f.scalar("dpsi");
        t.pres_scale = f.scalar("pres_scale");
        t.ne_scale = f.scalar("ne_scale");
        t.te_scale = f.scalar("te_scale");
        t.num_cols = f.dim("numz");
        for (size_t i = 0; i < 4; i++) {
            const std::string n = std::to_string(i);
            t.fpol[i] = f.get("fpol_c" + n);
            t.pres[i] = f.get("pressure_c" + n);
            t.te[i] = f.get("te_c" + n);
            t.ne[i] = f.get("ne_c" + n);
            for (size_t j = 0; j < 4; j++) t.psi[i][j] = f.get("psi_c" + n + std::to_string(j));
        }
//  Reference constructor quirk: ne_c0(te_c0), ne_c1(te_c1)  (equilibrium.hpp:1478).
        t.ne[0] = t.te[0];
        t.ne[1] = t.te[1];
        return t;
    }

    template<typename T=double, bool SAFE_MATH=false>
    shared<T, SAFE_MATH> make_efit(const std::string &spline_file) {
        return std::make_shared<efit<T, SAFE_MATH>> (load_efit_tables(spline_file));
    }
    template<typename T=double, bool SAFE_MATH=false>
    shared<T, SAFE_MATH> make_efit(const efit<>::tables &t) {
        return std::make_shared<efit<T, SAFE_MATH>> (t);
    }
//------------------------------------------------------------------------------
///  VMEC stellarator equilibrium in flux coordinates (s, u, v)  (equilibrium.hpp:1868-2413).
///  R, Z, lambda are Fourier series in (m u - n v) whose amplitudes are radial cubic splines;
///  the contravariant basis, the Jacobian and B = (J B^u e_u + J B^v e_v)/J follow by df().
///  Reference quirk kept: chi is evaluated with the NORMALISED s as spline argument
///  (equilibrium.hpp:2138 vs :2045-2050).
//------------------------------------------------------------------------------
    template<typename T=double, bool SAFE_MATH=false>
    class vmec final : public generic<T, SAFE_MATH> {
    public:
        struct tables {
            double sminh, sminf, ds, dphi, signj;
            std::array<std::vector<double>, 4> chi;
            std::array<std::vector<std::vector<double>>, 4> rmnc, zmns, lmns;     // [coefficient][mode][radial cell]
            std::vector<double> xm, xn;
        };
    private:
        const tables tab;
        graph::table_ptr rmnc_table, zmns_table, lmns_table;
        leaf_ptr s_cache, u_cache, v_cache, x_cache, y_cache, z_cache;
        vector_ptr esups_cache, esupu_cache, esupv_cache, bvec_cache;

        leaf_ptr spline(const std::array<std::vector<std::vector<double>>, 4> &c, const size_t mode,
                        leaf_ptr s, const double offset) {
            return build_1D_spline({graph::piecewise_1D(c[0][mode], s, tab.ds, offset),
                                    graph::piecewise_1D(c[1][mode], s, tab.ds, offset),
                                    graph::piecewise_1D(c[2][mode], s, tab.ds, offset),
                                    graph::piecewise_1D(c[3][mode], s, tab.ds, offset)}, s, tab.ds, offset);
        }
///  Cylindrical (dR, R dphi, dZ) components rotated to Cartesian (equilibrium.hpp:1948-2010).
        vector_ptr rotate(leaf_ptr a, leaf_ptr b, leaf_ptr c) {
            auto cosv = graph::cos(v_cache);
            auto sinv = graph::sin(v_cache);
            auto zero = graph::zero();
            auto m = graph::matrix(graph::vector(cosv, -sinv, zero),
                                   graph::vector(sinv, cosv, zero),
                                   graph::vector(zero, zero, graph::one()));
            return m->dot(graph::vector(a, b, c));
        }
        leaf_ptr get_chi(leaf_ptr s) {
            return build_1D_spline({graph::piecewise_1D(tab.chi[0], s, tab.ds, tab.sminf),
                                    graph::piecewise_1D(tab.chi[1], s, tab.ds, tab.sminf),
                                    graph::piecewise_1D(tab.chi[2], s, tab.ds, tab.sminf),
                                    graph::piecewise_1D(tab.chi[3], s, tab.ds, tab.sminf)}, s, tab.ds, tab.sminf);
        }
        void set_cache(leaf_ptr s, leaf_ptr u, leaf_ptr v) {
            if (s->is_match(s_cache) && u->is_match(u_cache) && v->is_match(v_cache)) return;
            s_cache = s;
            u_cache = u;
            v_cache = v;
            auto s_norm_f = (s - tab.sminf)/tab.ds;
            auto zero = graph::zero();
            leaf_ptr r = zero, z = zero, l = zero;
            if (use_mode_loop()) {
//  R, Z, lambda as Fourier-series nodes: their derivatives stay in the family and the emitter
//  evaluates all of them in one device loop over the 86 modes (graph::fourier_series).
                r = graph::fourier_series(rmnc_table, s, u, v, tab.ds, tab.sminf, {0, 0, 0, 0});
                z = graph::fourier_series(zmns_table, s, u, v, tab.ds, tab.sminf, {0, 0, 0, 1});
                l = graph::fourier_series(lmns_table, s, u, v, tab.ds, tab.sminh, {0, 0, 0, 1});
            } else {
//  The reference's construction: every mode unrolled in the graph (equilibrium.hpp:2120-2151).
                for (size_t i = 0, ie = tab.xm.size(); i < ie; i++) {
                    auto rmnc = spline(tab.rmnc, i, s, tab.sminf);
                    auto zmns = spline(tab.zmns, i, s, tab.sminf);
                    auto lmns = spline(tab.lmns, i, s, tab.sminh);
                    auto angle = graph::constant(tab.xm[i])*u - graph::constant(tab.xn[i])*v;
                    auto sinmn = graph::sin(angle);
                    r = r + rmnc*graph::cos(angle);
                    z = z + zmns*sinmn;
                    l = l + lmns*sinmn;
                }
            }
            x_cache = r*graph::cos(v);
            y_cache = r*graph::sin(v);
            z_cache = z;
            auto esubs = rotate(r->df(s), zero, z->df(s));
            auto esubu = rotate(r->df(u), zero, z->df(u));
            auto esubv = rotate(r->df(v), r, z->df(v));
            auto jacobian = esubs->dot(esubu->cross(esubv));
            esups_cache = esubu->cross(esubv)/jacobian;
            esupu_cache = esubv->cross(esubs)/jacobian;
            esupv_cache = esubs->cross(esubu)/jacobian;
            auto phip = (graph::constant(tab.signj)*graph::constant(tab.dphi)*s)->df(s);
            auto jbsupu = get_chi(s_norm_f)->df(s) - phip*l->df(v);
            auto jbsupv = phip*(1.0 + l->df(u));
            bvec_cache = (jbsupu*esubu + jbsupv*esubv)/jacobian;
        }
///  (1 - |s|^1.5)^2  (equilibrium.hpp:2160-2163).
        leaf_ptr get_profile(leaf_ptr s) {
            return graph::pow(1.0 - graph::pow(graph::sqrt(s*s), 1.5), 2.0);
        }
    public:
        vmec(const tables &t) : generic<T, SAFE_MATH> ({deuterium_mass}, {1}), tab(t) {
            s_cache = u_cache = v_cache = graph::zero();
            rmnc_table = graph::fourier_table(tab.rmnc, tab.xm, tab.xn, tab.ds, tab.sminf);
            zmns_table = graph::fourier_table(tab.zmns, tab.xm, tab.xn, tab.ds, tab.sminf);
            lmns_table = graph::fourier_table(tab.lmns, tab.xm, tab.xn, tab.ds, tab.sminh);
        }
///  true (default): Fourier sums are device loops; false: the reference's fully unrolled graph
///  (kept for cross-checks; 9 k statements per right-hand side).
        static bool &use_mode_loop() {
            static thread_local bool on = std::getenv("GFB_VMEC_UNROLLED") == nullptr;
            return on;
        }
        virtual vector_ptr get_esup1(leaf_ptr s, leaf_ptr u, leaf_ptr v) { set_cache(s, u, v); return esups_cache; }
        virtual vector_ptr get_esup2(leaf_ptr s, leaf_ptr u, leaf_ptr v) { set_cache(s, u, v); return esupu_cache; }
        virtual vector_ptr get_esup3(leaf_ptr s, leaf_ptr u, leaf_ptr v) { set_cache(s, u, v); return esupv_cache; }
        virtual leaf_ptr get_electron_density(leaf_ptr s, leaf_ptr, leaf_ptr) { return graph::constant(1.0E19)*get_profile(s); }
        virtual leaf_ptr get_ion_density(const size_t, leaf_ptr s, leaf_ptr u, leaf_ptr v) { return get_electron_density(s, u, v); }
        virtual leaf_ptr get_electron_temperature(leaf_ptr s, leaf_ptr, leaf_ptr) { return graph::constant(1000.0)*get_profile(s); }
        virtual leaf_ptr get_ion_temperature(const size_t, leaf_ptr s, leaf_ptr u, leaf_ptr v) { return get_electron_temperature(s, u, v); }
        virtual vector_ptr get_magnetic_field(leaf_ptr s, leaf_ptr u, leaf_ptr v) { set_cache(s, u, v); return bvec_cache; }
        virtual leaf_ptr get_characteristic_field(const size_t=0) final {
            auto zero = graph::zero();
            return get_magnetic_field(zero, zero, zero)->length();
        }
///  The radial full grid of the Fourier amplitude splines (first coordinate s).
        virtual cell_grid get_cell_grid() const {
            cell_grid g;
            const size_t cells = tab.rmnc[0].empty() ? 0 : tab.rmnc[0][0].size();
            g.dims = cells ? 1 : 0;
            g.lo[0] = tab.sminf; g.hi[0] = tab.sminf + tab.ds*static_cast<double> (cells); g.cells[0] = static_cast<unsigned> (cells);
            g.cell_length = 2.0*tab.ds;         // s is normalised flux: a cell is about 2 ds of the minor radius
            return g;
        }
        virtual leaf_ptr get_x(leaf_ptr s, leaf_ptr u, leaf_ptr v) { set_cache(s, u, v); return x_cache; }
        virtual leaf_ptr get_y(leaf_ptr s, leaf_ptr u, leaf_ptr v) { set_cache(s, u, v); return y_cache; }
        virtual leaf_ptr get_z(leaf_ptr s, leaf_ptr u, leaf_ptr v) { set_cache(s, u, v); return z_cache; }
    };

///  Load VMEC tables (variable names of equilibrium.hpp:2424-2640) from a GFBT file.
    inline vmec<>::tables load_vmec_tables(const std::string &spline_file) {
        table_file f(spline_file);
        vmec<>::tables t;
        t.sminf = f.scalar("sminf"); t.sminh = f.scalar("sminh"); t.ds = f.scalar("ds");
        t.dphi = f.scalar("dphi"); t.signj = f.scalar("signj");
        t.xm = f.get("xm");
        t.xn = f.get("xn");
        const size_t modes = t.xm.size();
        auto rows = [modes] (const std::vector<double> &flat) {
            const size_t cols = flat.size()/modes;
            std::vector<std::vector<double>> r(modes);
            for (size_t i = 0; i < modes; i++) r[i].assign(flat.begin() + i*cols, flat.begin() + (i + 1)*cols);
            return r;
        };
        for (size_t i = 0; i < 4; i++) {
            const std::string n = std::to_string(i);
            t.chi[i] = f.get("chi_c" + n);
            t.rmnc[i] = rows(f.get("rmnc_c" + n));
            t.zmns[i] = rows(f.get("zmns_c" + n));
            t.lmns[i] = rows(f.get("lmns_c" + n));
        }
        return t;
    }
    template<typename T=double, bool SAFE_MATH=false>
    shared<T, SAFE_MATH> make_vmec(const std::string &spline_file) {
        return std::make_shared<vmec<T, SAFE_MATH>> (load_vmec_tables(spline_file));
    }
}

#endif /* gfb_graph_equilibrium_hpp */
