//------------------------------------------------------------------------------
//  vector.hpp -- 3-vectors and 3x3 matrices of graph nodes.
//  Mirrors /root/reference/graph_framework/vector.hpp (vector_quantity: dot,
//  cross, length, unit, df, remove_pseudo; matrix_quantity::dot).
//------------------------------------------------------------------------------
#ifndef gfb_graph_vector_hpp
#define gfb_graph_vector_hpp

#include "node.hpp"

namespace graph {
    class vector_quantity;
    using vector_ptr = std::shared_ptr<vector_quantity>;
    template<typename T=double, bool SAFE_MATH=false>
    using shared_vector = vector_ptr;

    class vector_quantity : public std::enable_shared_from_this<vector_quantity> {
    protected:
        leaf_ptr x, y, z;
    public:
        vector_quantity(leaf_ptr x, leaf_ptr y, leaf_ptr z) : x(x), y(y), z(z) {}
        leaf_ptr get_x() const { return x; }
        leaf_ptr get_y() const { return y; }
        leaf_ptr get_z() const { return z; }
        leaf_ptr dot(vector_ptr v) { return x*v->get_x() + y*v->get_y() + z*v->get_z(); }
        vector_ptr cross(vector_ptr v) {
            return std::make_shared<vector_quantity> (y*v->get_z() - z*v->get_y(),
                                                      z*v->get_x() - x*v->get_z(),
                                                      x*v->get_y() - y*v->get_x());
        }
        leaf_ptr length() { return sqrt(dot(shared_from_this())); }
        vector_ptr unit() {
            auto l = length();
            return std::make_shared<vector_quantity> (x/l, y/l, z/l);
        }
        vector_ptr df(leaf_ptr arg) {
            return std::make_shared<vector_quantity> (x->df(arg), y->df(arg), z->df(arg));
        }
        vector_ptr remove_pseudo() {
            return std::make_shared<vector_quantity> (x->remove_pseudo(), y->remove_pseudo(), z->remove_pseudo());
        }
    };

    inline vector_ptr vector(leaf_ptr x, leaf_ptr y, leaf_ptr z) { return std::make_shared<vector_quantity> (x, y, z); }
    inline vector_ptr vector(const double x, leaf_ptr y, leaf_ptr z) { return vector(constant(x), y, z); }
    inline vector_ptr vector(leaf_ptr x, const double y, leaf_ptr z) { return vector(x, constant(y), z); }
    inline vector_ptr vector(leaf_ptr x, leaf_ptr y, const double z) { return vector(x, y, constant(z)); }
    inline vector_ptr vector(const double x, const double y, leaf_ptr z) { return vector(constant(x), constant(y), z); }
    inline vector_ptr vector(leaf_ptr x, const double y, const double z) { return vector(x, constant(y), constant(z)); }
    inline vector_ptr vector(const double x, leaf_ptr y, const double z) { return vector(constant(x), y, constant(z)); }
    inline vector_ptr vector(const double x, const double y, const double z) {
        return vector(constant(x), constant(y), constant(z));
    }

    inline vector_ptr operator+(vector_ptr l, vector_ptr r) {
        return vector(l->get_x() + r->get_x(), l->get_y() + r->get_y(), l->get_z() + r->get_z());
    }
    inline vector_ptr operator-(vector_ptr l, vector_ptr r) {
        return vector(l->get_x() - r->get_x(), l->get_y() - r->get_y(), l->get_z() - r->get_z());
    }
    inline vector_ptr operator*(leaf_ptr s, vector_ptr v) { return vector(s*v->get_x(), s*v->get_y(), s*v->get_z()); }
    inline vector_ptr operator*(const double s, vector_ptr v) { return constant(s)*v; }
    inline vector_ptr operator/(vector_ptr v, leaf_ptr s) { return vector(v->get_x()/s, v->get_y()/s, v->get_z()/s); }
    inline vector_ptr operator/(vector_ptr v, const double s) { return v/constant(s); }

    class matrix_quantity {
    protected:
        vector_ptr r1, r2, r3;
    public:
        matrix_quantity(vector_ptr r1, vector_ptr r2, vector_ptr r3) : r1(r1), r2(r2), r3(r3) {}
        vector_ptr dot(vector_ptr v) { return vector(r1->dot(v), r2->dot(v), r3->dot(v)); }
    };
    template<typename T=double, bool SAFE_MATH=false>
    using shared_matrix = std::shared_ptr<matrix_quantity>;
    inline std::shared_ptr<matrix_quantity> matrix(vector_ptr r1, vector_ptr r2, vector_ptr r3) {
        return std::make_shared<matrix_quantity> (r1, r2, r3);
    }
}

#endif /* gfb_graph_vector_hpp */
