// Host-side quadtree builder behind efgpu_mesh_* (include/efgpu.h).
//
// Reproduces, for the single-tree square connectivity the reference uses (src/P4est.cpp:7-33),
// what Mesh::refineByFunction (src/Mesh.hpp:111-180) obtains from p4est:
//   p4est_new_ext(min_level, uniform) -> recursive p4est_refine with the reference's callback
//   (any of the nx*ny cell centres of the quadrant flags it, :134-163) -> p4est_balance(CORNER)
// and lays the nodes out the way the Quadtree constructor visits them (src/Quadtree.hpp:118-196):
// depth-first, children in Morton order, child boxes by midpoint splitting of the parent box
// (FiniteVolumeNodeFactory.cpp:24-57).  The 2:1 corner-balanced refinement of a given leaf set is
// unique, so the result is identical to p4est's (pinned against reference dumps in tests/).
#include "../../include/efgpu.h"

#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_set>
#include <vector>

namespace {

constexpr int kMaxLevel = 30;                       // P4EST_MAXLEVEL
constexpr double kRootLen = (double)(1 << kMaxLevel);  // P4EST_ROOT_LEN

struct Quad { int l, x, y; };
inline uint64_t key(int l, int x, int y) { return ((uint64_t)l << 56) | ((uint64_t)x << 28) | (uint64_t)y; }

// p4est_qcoord_to_vertex (extern/p4est/src/p4est.c:82-143) for the four corner vertices
inline void qcoord_to_vertex(const double box[4], double qx, double qy, double* vx, double* vy)
{
    const double wx1 = qx / kRootLen, wx0 = 1.0 - wx1, wy1 = qy / kRootLen, wy0 = 1.0 - wy1;
    const double cx[4] = {box[0], box[1], box[0], box[1]}, cy[4] = {box[2], box[2], box[3], box[3]};
    const double wy[4] = {wy0, wy0, wy1, wy1}, wx[4] = {wx0, wx1, wx0, wx1};
    double x = 0.0, y = 0.0;
    for (int v = 0; v < 4; v++) { const double f = wy[v] * wx[v]; x += f * cx[v]; y += f * cy[v]; }
    *vx = x; *vy = y;
}

}  // namespace

struct efgpu_mesh {
    int nx = 0;
    std::vector<int32_t> level, child, leaf_nodes;
    std::vector<double> box;
    std::vector<std::string> path;
};

extern "C" {

int efgpu_refine_elliptic_single(double x, double y, void* user)
{
    return std::fabs(-(std::sin(x) + std::sin(y))) > *static_cast<const double*>(user) ? 1 : 0;
}

int efgpu_mesh_create(double xl, double xu, double yl, double yu, int nx, int min_level, int max_level,
                      efgpu_refine_fn fn, void* user, efgpu_mesh** out)
{
    if (!out || nx <= 0 || min_level < 0 || max_level < min_level || max_level > 20 || !(xl < xu) || !(yl < yu)) return EFGPU_ERR_BAD_ARG;
    const double root_box[4] = {xl, xu, yl, yu};
    std::unordered_set<uint64_t> leaves;
    std::vector<Quad> stack;
    for (int x = 0; x < (1 << min_level); x++)
        for (int y = 0; y < (1 << min_level); y++) stack.push_back({min_level, x, y});
    auto want_refine = [&](const Quad& q) {
        if (!fn || q.l >= max_level) return false;
        const double qlen = kRootLen / (double)(1 << q.l);
        double xlo, ylo, xhi, yhi;
        qcoord_to_vertex(root_box, q.x * qlen, q.y * qlen, &xlo, &ylo);
        qcoord_to_vertex(root_box, (q.x + 1) * qlen, (q.y + 1) * qlen, &xhi, &yhi);
        const double dx = (xhi - xlo) / nx, dy = (yhi - ylo) / nx;
        for (int i = 0; i < nx; i++)
            for (int j = 0; j < nx; j++)
                if (fn(xlo + (i + 0.5) * dx, ylo + (j + 0.5) * dy, user)) return true;
        return false;
    };
    while (!stack.empty()) {
        Quad q = stack.back(); stack.pop_back();
        if (want_refine(q)) { for (int c = 0; c < 4; c++) stack.push_back({q.l + 1, 2 * q.x + (c & 1), 2 * q.y + (c >> 1)}); }
        else leaves.insert(key(q.l, q.x, q.y));
    }
    // 2:1 balance across faces and corners
    auto containing_leaf = [&](int l, int x, int y, Quad* found) {
        for (; l >= 0; l--, x >>= 1, y >>= 1)
            if (leaves.count(key(l, x, y))) { *found = {l, x, y}; return true; }
        return false;
    };
    std::vector<Quad> work;
    for (uint64_t k : leaves) work.push_back({(int)(k >> 56), (int)((k >> 28) & 0xfffffff), (int)(k & 0xfffffff)});
    while (!work.empty()) {
        Quad q = work.back(); work.pop_back();
        if (!leaves.count(key(q.l, q.x, q.y))) continue;
        for (int dx = -1; dx <= 1; dx++)
            for (int dy = -1; dy <= 1; dy++) {
                if (!dx && !dy) continue;
                const int nxq = q.x + dx, nyq = q.y + dy;
                if (nxq < 0 || nyq < 0 || nxq >= (1 << q.l) || nyq >= (1 << q.l)) continue;
                Quad nb;
                while (containing_leaf(q.l, nxq, nyq, &nb) && nb.l < q.l - 1) {
                    leaves.erase(key(nb.l, nb.x, nb.y));
                    for (int c = 0; c < 4; c++) {
                        Quad k{nb.l + 1, 2 * nb.x + (c & 1), 2 * nb.y + (c >> 1)};
                        leaves.insert(key(k.l, k.x, k.y)); work.push_back(k);
                    }
                }
            }
    }
    // node table, depth first, children in Morton order
    efgpu_mesh* m = new efgpu_mesh();
    m->nx = nx;
    struct Frame { Quad q; int parent, cid; double box[4]; std::string path; };
    std::vector<Frame> st;
    st.push_back({{0, 0, 0}, -1, 0, {xl, xu, yl, yu}, "0"});
    while (!st.empty()) {
        Frame f = st.back(); st.pop_back();
        const int id = (int)m->level.size();
        m->level.push_back(f.q.l);
        for (int c = 0; c < 4; c++) m->child.push_back(-1);
        for (int c = 0; c < 4; c++) m->box.push_back(f.box[c]);
        m->path.push_back(f.path);
        if (f.parent >= 0) m->child[4 * (size_t)f.parent + f.cid] = id;
        if (leaves.count(key(f.q.l, f.q.x, f.q.y))) { m->leaf_nodes.push_back(id); continue; }
        const double xm = (f.box[0] + f.box[1]) / 2.0, ym = (f.box[2] + f.box[3]) / 2.0;
        const double cb[4][4] = {{f.box[0], xm, f.box[2], ym}, {xm, f.box[1], f.box[2], ym}, {f.box[0], xm, ym, f.box[3]}, {xm, f.box[1], ym, f.box[3]}};
        for (int c = 3; c >= 0; c--) {   // pushed in reverse so child 0 is visited first
            Frame k{{f.q.l + 1, 2 * f.q.x + (c & 1), 2 * f.q.y + (c >> 1)}, id, c, {cb[c][0], cb[c][1], cb[c][2], cb[c][3]}, f.path + (char)('0' + c)};
            st.push_back(k);
        }
    }
    *out = m;
    return EFGPU_OK;
}

int efgpu_mesh_desc(const efgpu_mesh* m, efgpu_tree_desc* out)
{
    if (!m || !out) return EFGPU_ERR_BAD_ARG;
    out->n_nodes = (int32_t)m->level.size(); out->nx = m->nx;
    out->level = m->level.data(); out->child = m->child.data(); out->box = m->box.data();
    return EFGPU_OK;
}
int efgpu_mesh_n_leaves(const efgpu_mesh* m) { return m ? (int)m->leaf_nodes.size() : -1; }
const int32_t* efgpu_mesh_leaf_nodes(const efgpu_mesh* m) { return m ? m->leaf_nodes.data() : nullptr; }
int efgpu_mesh_path(const efgpu_mesh* m, int node, char* buf, size_t capacity)
{
    if (!m || !buf || node < 0 || node >= (int)m->path.size() || capacity <= m->path[node].size()) return EFGPU_ERR_BAD_ARG;
    std::memcpy(buf, m->path[node].c_str(), m->path[node].size() + 1);
    return EFGPU_OK;
}
void efgpu_mesh_destroy(efgpu_mesh* m) { delete m; }

}  // extern "C"
