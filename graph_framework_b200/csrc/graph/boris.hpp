//------------------------------------------------------------------------------
//  boris.hpp -- the relativistic Boris push of /root/reference/graph_korc/xkorc.cpp:40-121 as a
//  graph: u' = u - dt u x B/(2 gamma), tau = -dt B/2, gamma+ = sqrt((sigma + sqrt(sigma^2 +
//  4 (tau^2 + (u'.tau)^2)))/2), rotation, x += rho_L dt u+/gamma+, with B normalised by the on-axis
//  field b0 and lengths by the Larmor radius c m_e/(q b0).  Shared by the C binding (gfb_boris_*) and the
//  CPU emission tests.
//------------------------------------------------------------------------------
#ifndef gfb_graph_boris_hpp
#define gfb_graph_boris_hpp

#include "equilibrium.hpp"

namespace boris {
    using graph::leaf_ptr;

    struct push_graph {
        graph::map_nodes<> initialize;      ///< the `initialize_gamma` pre-item: u <- gamma u, gamma
        graph::map_nodes<> step;            ///< the `step` item
        double larmor_radius;
    };

///  vars = x, y, z, ux, uy, uz, gamma; b0 = |B| on the magnetic axis (equilibrium::get_characteristic_field).
    inline push_graph build(equilibrium::shared<> &eq, const std::vector<leaf_ptr> &vars, leaf_ptr b0, const double dt_value) {
        const double q = 1.602176634E-19;
        const double me = 9.1093837139E-31;
        const double c = 299792458.0;
        auto gryo_period = me/(q*b0);
        auto larmor_radius = c*gryo_period;
        auto x = vars[0], y = vars[1], z = vars[2], ux = vars[3], uy = vars[4], uz = vars[5], gamma = vars[6];
        auto pos = graph::vector(x, y, z);
        auto u_vec = graph::vector(ux, uy, uz);
        auto dt = graph::constant(dt_value);
//  xkorc.cpp:66-121
        auto gamma_init = 1.0/graph::sqrt(1.0 - u_vec->dot(u_vec));
        auto u_init = gamma_init*u_vec;
        auto b_vec = eq->get_magnetic_field(x, y, z)/b0;

        auto u_prime = u_vec - dt*u_vec->cross(b_vec)/(2.0*gamma);
        auto tau = -0.5*dt*b_vec;
        auto tau_sq = tau->dot(tau);
        auto speed_sq = u_prime->dot(u_prime);
        auto sigma = 1.0 + speed_sq - tau_sq;
        auto ustar = u_prime->dot(tau);
        auto gamma_next = graph::sqrt(0.5*(sigma + graph::sqrt(sigma*sigma + 4.0*(tau_sq + ustar*ustar))));
        auto t = tau/gamma_next;
        auto s = 1.0 + t->dot(t);
        auto u_prime_dot_t = u_prime->dot(t);
        auto u_next = (u_prime + u_prime_dot_t*t + u_prime->cross(t))/s;
        auto pos_next = pos + larmor_radius*dt*u_next/gamma_next;

        push_graph g;
        g.initialize = {{u_init->get_x(), ux}, {u_init->get_y(), uy}, {u_init->get_z(), uz}, {gamma_init, gamma}};
        g.step = {{pos_next->get_x(), x}, {pos_next->get_y(), y}, {pos_next->get_z(), z},
                  {u_next->get_x(), ux}, {u_next->get_y(), uy}, {u_next->get_z(), uz}, {gamma_next, gamma}};
        g.larmor_radius = larmor_radius->evaluate().at(0);
        return g;
    }
}

#endif /* gfb_graph_boris_hpp */
