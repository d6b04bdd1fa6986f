// netcdf.h stand-in used ONLY to build the reference oracle (oracle/_ref).
// TEST INFRASTRUCTURE -- not product code.
//
// The reference links libnetcdf (equilibrium.hpp:214, output.hpp:11); this image
// has none.  Reads are served from a GFBT table file (see
// graph_framework_b200/tools/gfbt.py); every define/write call is a no-op so
// that solver_interface's result_file (solver.hpp:231-234) costs nothing.
#ifndef GFB_ORACLE_NETCDF_SHIM_H
#define GFB_ORACLE_NETCDF_SHIM_H

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

typedef int nc_type;
#define NC_NOERR 0
#define NC_NOWRITE 0
#define NC_WRITE 1
#define NC_CLOBBER 0
#define NC_DISKLESS 8
#define NC_FLOAT 5
#define NC_DOUBLE 6
#define NC_UNLIMITED 0L

namespace gfb_nc_shim {
struct var { std::vector<size_t> dims; std::vector<double> data; };
struct file {
    std::vector<std::string> names;
    std::vector<var> vars;
    std::vector<std::string> dim_names;
    std::vector<size_t> dim_lens;
};
inline std::mutex &lock() { static std::mutex m; return m; }
inline std::map<int, file> &table() { static std::map<int, file> t; return t; }
inline int &next_id() { static int n = 1; return n; }

inline bool load(const char *path, file &f) {
    FILE *fp = std::fopen(path, "rb");
    if (!fp) return false;
    char magic[6];
    if (std::fread(magic, 1, 6, fp) != 6 || std::memcmp(magic, "GFBT1\n", 6)) { std::fclose(fp); return false; }
    uint32_t n = 0;
    if (std::fread(&n, 4, 1, fp) != 1) { std::fclose(fp); return false; }
    for (uint32_t i = 0; i < n; i++) {
        uint32_t ln = 0, rank = 0;
        if (std::fread(&ln, 4, 1, fp) != 1) break;
        std::string name(ln, '\0');
        if (std::fread(name.data(), 1, ln, fp) != ln) break;
        if (std::fread(&rank, 4, 1, fp) != 1) break;
        var v; size_t cnt = 1;
        for (uint32_t d = 0; d < rank; d++) { uint64_t e; if (std::fread(&e, 8, 1, fp) != 1) break; v.dims.push_back(e); cnt *= e; }
        v.data.resize(cnt);
        if (std::fread(v.data.data(), 8, cnt, fp) != cnt) break;
        if (name.rfind("dim:", 0) == 0) {
            f.dim_names.push_back(name.substr(4));
            f.dim_lens.push_back(static_cast<size_t>(v.data[0]));
        } else {
            f.names.push_back(name);
            f.vars.push_back(std::move(v));
        }
    }
    std::fclose(fp);
    return true;
}
}  // namespace gfb_nc_shim

inline const char *nc_strerror(int) { return "netcdf shim"; }
inline int nc_open(const char *path, int, int *ncid) {
    std::lock_guard<std::mutex> g(gfb_nc_shim::lock());
    gfb_nc_shim::file f;
    if (!gfb_nc_shim::load(path, f)) { std::fprintf(stderr, "netcdf shim: cannot read %s\n", path); return 1; }
    *ncid = gfb_nc_shim::next_id()++;
    gfb_nc_shim::table()[*ncid] = std::move(f);
    return NC_NOERR;
}
inline int nc_close(int ncid) {
    std::lock_guard<std::mutex> g(gfb_nc_shim::lock());
    gfb_nc_shim::table().erase(ncid);
    return NC_NOERR;
}
inline int nc_inq_varid(int ncid, const char *name, int *varid) {
    std::lock_guard<std::mutex> g(gfb_nc_shim::lock());
    auto it = gfb_nc_shim::table().find(ncid);
    if (it == gfb_nc_shim::table().end()) { *varid = -1; return 1; }
    auto &f = it->second;
    for (size_t i = 0; i < f.names.size(); i++) if (f.names[i] == name) { *varid = int(i); return NC_NOERR; }
    *varid = -1;
    return 1;
}
inline int nc_inq_dimid(int ncid, const char *name, int *dimid) {
    std::lock_guard<std::mutex> g(gfb_nc_shim::lock());
    auto it = gfb_nc_shim::table().find(ncid);
    if (it == gfb_nc_shim::table().end()) { *dimid = -1; return 1; }
    auto &f = it->second;
    for (size_t i = 0; i < f.dim_names.size(); i++) if (f.dim_names[i] == name) { *dimid = int(i); return NC_NOERR; }
    *dimid = -1;
    return 1;
}
inline int nc_inq_dimlen(int ncid, int dimid, size_t *len) {
    std::lock_guard<std::mutex> g(gfb_nc_shim::lock());
    auto it = gfb_nc_shim::table().find(ncid);
    if (it == gfb_nc_shim::table().end() || dimid < 0) { *len = 0; return 1; }
    *len = it->second.dim_lens[dimid];
    return NC_NOERR;
}
template<typename T> inline int nc_get_var(int ncid, int varid, T *out) {
    std::lock_guard<std::mutex> g(gfb_nc_shim::lock());
    auto it = gfb_nc_shim::table().find(ncid);
    if (it == gfb_nc_shim::table().end() || varid < 0) return 1;
    auto &v = it->second.vars[varid];
    for (size_t i = 0; i < v.data.size(); i++) out[i] = static_cast<T>(v.data[i]);
    return NC_NOERR;
}
template<typename T> inline int nc_get_vara(int ncid, int varid, const size_t *start, const size_t *count, T *out) {
    std::lock_guard<std::mutex> g(gfb_nc_shim::lock());
    auto it = gfb_nc_shim::table().find(ncid);
    if (it == gfb_nc_shim::table().end() || varid < 0) return 1;
    auto &v = it->second.vars[varid];
    const size_t rank = v.dims.size();
    if (rank == 1) {
        for (size_t i = 0; i < count[0]; i++) out[i] = static_cast<T>(v.data[start[0] + i]);
    } else if (rank == 2) {
        size_t k = 0;
        for (size_t i = 0; i < count[0]; i++)
            for (size_t j = 0; j < count[1]; j++)
                out[k++] = static_cast<T>(v.data[(start[0] + i)*v.dims[1] + start[1] + j]);
    } else {
        return 1;
    }
    return NC_NOERR;
}
// ---- result-file side: nothing is written by the oracle ---------------------
inline int nc_create(const char *, int, int *ncid) { *ncid = 0; return NC_NOERR; }
inline int nc_def_dim(int, const char *, size_t, int *dimid) { *dimid = 0; return NC_NOERR; }
inline int nc_def_var(int, const char *, nc_type, int, const int *, int *varid) { *varid = 0; return NC_NOERR; }
inline int nc_enddef(int) { return NC_NOERR; }
inline int nc_redef(int) { return NC_NOERR; }
inline int nc_sync(int) { return NC_NOERR; }
inline int nc_inq_var(int, int, char *, nc_type *type, int *ndims, int *, int *) { if (type) *type = NC_DOUBLE; if (ndims) *ndims = 0; return 1; }
inline int nc_put_vara_double(int, int, const size_t *, const size_t *, const double *) { return NC_NOERR; }
inline int nc_put_vara_float(int, int, const size_t *, const size_t *, const float *) { return NC_NOERR; }
inline int nc_get_varm_double(int, int, const size_t *, const size_t *, const ptrdiff_t *, const ptrdiff_t *, double *) { return 1; }
inline int nc_get_varm_float(int, int, const size_t *, const size_t *, const ptrdiff_t *, const ptrdiff_t *, float *) { return 1; }

#endif
