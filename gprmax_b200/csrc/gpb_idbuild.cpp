// gpb_idbuild.cpp -- host-side build of the per-edge material ID array.
//
// Replaces the bulk of the reference's build_electric_components / build_magnetic_components
// (gprMax/yee_cell_build_ext.pyx:110-257): a single-threaded triple loop over the 6 N edges of the grid that, for every
// edge not marked rigid, looks at the 4 (electric) or 2 (magnetic) cells of `solid` around it and either copies their common
// material or asks for a dielectric-smoothed average (create_electric_average / create_magnetic_average, :31-107).
//
// The sequential part of that algorithm is tiny: which averaged material a combination of cell materials maps to depends on the
// ORDER in which distinct combinations are first met (new materials are appended in that order), not on how many edges show
// them.  So the work is split:
//   pass 1 (here, all host cores, planes split over threads): write every edge whose surrounding cells agree and collect the
//           distinct disagreeing combinations together with the first edge -- in the reference's scan order: component, i, j,
//           k -- that shows each of them;
//   host   (Python, gprmax_b200/yee_build.py): for the distinct combinations in that order call the reference's OWN
//           create_*_average on the recorded edge (a few hundred calls), which yields exactly the material numbering of the
//           reference;
//   pass 2 (here, all host cores): write the resolved material on every disagreeing edge (pass 1 marks them in ID, so pass 2
//           only streams through ID and looks at the cells of the marked edges again).
// The resulting ID array is bit-identical with the reference's (tests/test_yee_build.py).
#include "../../include/gprmax_b200.h"

#include <algorithm>
#include <cstring>
#include <map>
#include <thread>
#include <unistd.h>
#include <vector>

namespace {

struct Grid {
    const uint32_t *solid;   // [nx][ny][nz]
    const int8_t *rigidE;    // [12][nx][ny][nz]
    const int8_t *rigidH;    // [6][nx][ny][nz]
    uint32_t *ID;            // [6][nx+1][ny+1][nz+1]
    int nx, ny, nz;          // the whole domain
    // a slab-local build holds only a range of x planes of every array: cell planes [sx0, sx0 + snx) of solid / rigidE / rigidH,
    // node planes [ix0, ix0 + inx) of ID (whole arrays: 0, nx and 0, nx + 1); i, j, k below are always global indices
    int sx0, snx, ix0, inx;
    size_t cell(int i, int j, int k) const { return ((size_t)(i - sx0) * ny + j) * nz + k; }
    size_t node(int c, int i, int j, int k) const { return (((size_t)c * inx + (i - ix0)) * (ny + 1) + j) * (nz + 1) + k; }
    bool rE(int q, int i, int j, int k) const { return rigidE[(size_t)q * snx * ny * nz + cell(i, j, k)] != 0; }
    bool rH(int q, int i, int j, int k) const { return rigidH[(size_t)q * snx * ny * nz + cell(i, j, k)] != 0; }
    // planes [x0, x1) can be built from these arrays: the edges of node plane i look at the cell planes i - 1 and i
    bool covers(int x0, int x1) const
    {
        if (x1 <= x0) return true;
        return x0 >= ix0 && x1 <= ix0 + inx && sx0 <= std::max(x0 - 1, 0) && sx0 + snx >= std::min(x1, nx);
    }
    // yee_cell_setget_rigid_ext.pyx:24-66, 108-138 (loop ranges keep every index inside the arrays)
    bool rigid(int comp, int i, int j, int k) const
    {
        switch (comp) {
        case 0: return rE(0, i, j, k) || rE(1, i, j - 1, k) || rE(3, i, j, k - 1) || rE(2, i, j - 1, k - 1);
        case 1: return rE(4, i, j, k) || rE(7, i - 1, j, k) || rE(5, i, j, k - 1) || rE(6, i - 1, j, k - 1);
        case 2: return rE(8, i, j, k) || rE(9, i - 1, j, k) || rE(11, i, j - 1, k) || rE(10, i - 1, j - 1, k);
        case 3: return rH(0, i, j, k) || rH(1, i - 1, j, k);
        case 4: return rH(2, i, j, k) || rH(3, i, j - 1, k);
        default: return rH(4, i, j, k) || rH(5, i, j, k - 1);
        }
    }
    // the cells around an edge in the order the reference names them numID1..4 (yee_cell_build_ext.pyx:131-134, 153-156,
    // 175-178; 210-211, 230-231, 250-251); magnetic components fill two entries
    void around(int comp, int i, int j, int k, uint32_t id[4]) const
    {
        id[2] = id[3] = 0;
        switch (comp) {
        case 0: id[0] = solid[cell(i, j, k)]; id[1] = solid[cell(i, j - 1, k)]; id[2] = solid[cell(i, j - 1, k - 1)]; id[3] = solid[cell(i, j, k - 1)]; break;
        case 1: id[0] = solid[cell(i, j, k)]; id[1] = solid[cell(i - 1, j, k)]; id[2] = solid[cell(i - 1, j, k - 1)]; id[3] = solid[cell(i, j, k - 1)]; break;
        case 2: id[0] = solid[cell(i, j, k)]; id[1] = solid[cell(i - 1, j, k)]; id[2] = solid[cell(i - 1, j - 1, k)]; id[3] = solid[cell(i, j - 1, k)]; break;
        case 3: id[0] = solid[cell(i, j, k)]; id[1] = solid[cell(i - 1, j, k)]; break;
        case 4: id[0] = solid[cell(i, j, k)]; id[1] = solid[cell(i, j - 1, k)]; break;
        default: id[0] = solid[cell(i, j, k)]; id[1] = solid[cell(i, j, k - 1)]; break;
        }
    }
};

// loop ranges of the six components (yee_cell_build_ext.pyx:123-125, 145-147, 167-169, 202-204, 222-224, 242-244)
inline void ranges(const Grid &g, int comp, int lo[3], int hi[3])
{
    const int l[6][3] = {{0, 1, 1}, {1, 0, 1}, {1, 1, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int a = 0; a < 3; ++a) lo[a] = l[comp][a];
    hi[0] = g.nx; hi[1] = g.ny; hi[2] = g.nz;
}

struct Key {
    int comp;
    uint32_t id[4];
    bool operator<(const Key &o) const
    {
        if (comp != o.comp) return comp < o.comp;
        return std::lexicographical_compare(id, id + 4, o.id, o.id + 4);
    }
};
struct Pos {
    int i, j, k;
    bool before(const Pos &o) const { return i != o.i ? i < o.i : (j != o.j ? j < o.j : k < o.k); }
};

int host_threads()
{
    const long n = sysconf(_SC_NPROCESSORS_CONF);
    return (int)std::max(1l, std::min(64l, n));
}

template <typename F>
void parallel_planes(int x0, int x1, F f)
{
    const int nt = std::max(1, std::min(host_threads(), x1 - x0));
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) {
        const int a = x0 + (int)((long long)(x1 - x0) * t / nt), b = x0 + (int)((long long)(x1 - x0) * (t + 1) / nt);
        th.emplace_back([=] {
            cpu_set_t all;   // see ScopedFullAffinity in gpb_core.cu: the caller is usually pinned to one core by OpenMP
            CPU_ZERO(&all);
            for (int c = 0; c < CPU_SETSIZE; ++c) CPU_SET(c, &all);
            pthread_setaffinity_np(pthread_self(), sizeof all, &all);
            f(t, a, b);
        });
    }
    for (auto &t : th) t.join();
}

}  // namespace

extern "C" {

// Row-wise form of Grid::rigid / Grid::around: for one (component, i, j) the taps are pointers to the k rows of the rigid and
// solid arrays plus a k offset (0 or -1), so the inner loop over k is a handful of byte / word loads per edge.
struct Row {
    const int8_t *r[4];      // rigid taps (magnetic components have two: they are listed twice)
    const uint32_t *s[4];    // solid taps (likewise)
    int rk[4], sk[4];        // k offsets of the taps
    uint32_t *id;
    int nid;
    bool rigid(int k) const { return (r[0][k + rk[0]] | r[1][k + rk[1]] | r[2][k + rk[2]] | r[3][k + rk[3]]) != 0; }
    uint32_t cell(int t, int k) const { return s[t][k + sk[t]]; }
};

// (array 0 = rigidE, 1 = rigidH; q; di, dj, dk) and (di, dj, dk): the same taps as Grid::rigid and Grid::around
const int kRigidTap[6][4][4] = {
    {{0, 0, 0, 0}, {1, 0, -1, 0}, {3, 0, 0, -1}, {2, 0, -1, -1}},
    {{4, 0, 0, 0}, {7, -1, 0, 0}, {5, 0, 0, -1}, {6, -1, 0, -1}},
    {{8, 0, 0, 0}, {9, -1, 0, 0}, {11, 0, -1, 0}, {10, -1, -1, 0}},
    {{0, 0, 0, 0}, {1, -1, 0, 0}, {0, 0, 0, 0}, {1, -1, 0, 0}},
    {{2, 0, 0, 0}, {3, 0, -1, 0}, {2, 0, 0, 0}, {3, 0, -1, 0}},
    {{4, 0, 0, 0}, {5, 0, 0, -1}, {4, 0, 0, 0}, {5, 0, 0, -1}}};
const int kSolidTap[6][4][3] = {
    {{0, 0, 0}, {0, -1, 0}, {0, -1, -1}, {0, 0, -1}},
    {{0, 0, 0}, {-1, 0, 0}, {-1, 0, -1}, {0, 0, -1}},
    {{0, 0, 0}, {-1, 0, 0}, {-1, -1, 0}, {0, -1, 0}},
    {{0, 0, 0}, {-1, 0, 0}, {0, 0, 0}, {-1, 0, 0}},
    {{0, 0, 0}, {0, -1, 0}, {0, 0, 0}, {0, -1, 0}},
    {{0, 0, 0}, {0, 0, -1}, {0, 0, 0}, {0, 0, -1}}};

inline Row row_of(const Grid &g, int comp, int i, int j)
{
    Row w;
    const size_t cells = (size_t)g.snx * g.ny * g.nz;
    for (int t = 0; t < 4; ++t) {
        const int *rt = kRigidTap[comp][t];
        const int8_t *base = comp < 3 ? g.rigidE : g.rigidH;
        // (a k offset of -1 belongs to components whose loops start at k = 1)
        w.r[t] = base + (size_t)rt[0] * cells + ((size_t)(i + rt[1] - g.sx0) * g.ny + (j + rt[2])) * g.nz;
        w.rk[t] = rt[3];
        const int *st = kSolidTap[comp][t];
        w.s[t] = g.solid + ((size_t)(i + st[0] - g.sx0) * g.ny + (j + st[1])) * g.nz;
        w.sk[t] = st[2];
    }
    w.id = g.ID + (((size_t)comp * g.inx + (i - g.ix0)) * (g.ny + 1) + j) * (g.nz + 1);
    w.nid = comp < 3 ? 4 : 2;
    return w;
}

// An edge whose surrounding cells disagree is marked in ID between the two passes (material numbers stay far below this)
constexpr uint32_t kPending = 0xffffffffu;

static int ids_scan(const Grid &g, int x0, int x1, gpb_idcombo_t *combos, int max_combos, int *ncombos)
{
    const int nt = std::max(1, std::min(host_threads(), std::max(1, x1 - x0)));
    std::vector<std::map<Key, Pos>> found(nt);
    parallel_planes(x0, x1, [&](int t, int a, int b) {
        std::map<Key, Pos> &mine = found[t];
        Key last{-1, {0, 0, 0, 0}};   // runs of the same combination are common: skip the map for them
        for (int comp = 0; comp < 6; ++comp) {
            int lo[3], hi[3];
            ranges(g, comp, lo, hi);
            for (int i = std::max(a, lo[0]); i < std::min(b, hi[0]); ++i)
                for (int j = lo[1]; j < hi[1]; ++j) {
                    const Row w = row_of(g, comp, i, j);
                    // branch-free sweep of the row: agreeing edges get their material, disagreeing ones the mark, rigid ones keep
                    // what they have; almost every row ends here
                    unsigned any_pending = 0;
                    for (int k = lo[2]; k < hi[2]; ++k) {
                        const uint32_t c0 = w.cell(0, k), c1 = w.cell(1, k), c2 = w.cell(2, k), c3 = w.cell(3, k);
                        const unsigned same = (unsigned)(c0 == c1) & (unsigned)(c0 == c2) & (unsigned)(c0 == c3);
                        const unsigned rig = w.rigid(k) ? 1u : 0u;
                        const uint32_t v = same ? c0 : kPending;
                        w.id[k] = rig ? w.id[k] : v;
                        any_pending |= (rig ^ 1u) & (same ^ 1u);
                    }
                    if (!any_pending) continue;
                    for (int k = lo[2]; k < hi[2]; ++k) {
                        if (w.id[k] != kPending || w.rigid(k)) continue;
                        const uint32_t c0 = w.cell(0, k), c1 = w.cell(1, k), c2 = w.cell(2, k), c3 = w.cell(3, k);
                        const uint32_t id[4] = {c0, c1, w.nid == 4 ? c2 : 0u, w.nid == 4 ? c3 : 0u};
                        if (last.comp == comp && !memcmp(last.id, id, sizeof id)) continue;
                        last.comp = comp;
                        memcpy(last.id, id, sizeof id);
                        mine.insert({last, Pos{i, j, k}});   // keeps the first (scan order inside this thread's planes)
                    }
                }
        }
    });
    // threads own increasing plane ranges, so the first thread that met a combination met it first in scan order
    std::map<Key, Pos> all;
    for (int t = 0; t < nt; ++t)
        for (auto &kv : found[t]) all.insert(kv);
    std::vector<std::pair<Key, Pos>> order(all.begin(), all.end());
    std::sort(order.begin(), order.end(), [](const std::pair<Key, Pos> &x, const std::pair<Key, Pos> &y) {
        if (x.first.comp != y.first.comp) return x.first.comp < y.first.comp;
        return x.second.before(y.second);
    });
    *ncombos = (int)order.size();
    if ((int)order.size() > max_combos) return 2;
    for (size_t q = 0; q < order.size(); ++q) {
        gpb_idcombo_t &c = combos[q];
        c.comp = order[q].first.comp;
        memcpy(c.id, order[q].first.id, sizeof c.id);
        c.i = order[q].second.i; c.j = order[q].second.j; c.k = order[q].second.k;
    }
    return 0;
}

static int ids_apply(const Grid &g, int x0, int x1, const gpb_idcombo_t *combos, const uint32_t *numid, int ncombos)
{
    std::map<Key, uint32_t> table;
    for (int q = 0; q < ncombos; ++q) {
        Key k{combos[q].comp, {combos[q].id[0], combos[q].id[1], combos[q].id[2], combos[q].id[3]}};
        table[k] = numid[q];
    }
    int missing = 0;
    parallel_planes(x0, x1, [&](int, int a, int b) {
        Key last{-1, {0, 0, 0, 0}};
        uint32_t last_num = 0;
        for (int comp = 0; comp < 6; ++comp) {
            int lo[3], hi[3];
            ranges(g, comp, lo, hi);
            for (int i = std::max(a, lo[0]); i < std::min(b, hi[0]); ++i)
                for (int j = lo[1]; j < hi[1]; ++j) {
                    const Row w = row_of(g, comp, i, j);
                    unsigned any_pending = 0;
                    for (int k = lo[2]; k < hi[2]; ++k) any_pending |= (unsigned)(w.id[k] == kPending);
                    if (!any_pending) continue;
                    for (int k = lo[2]; k < hi[2]; ++k) {
                        if (w.id[k] != kPending) continue;   // only the edges pass 1 marked look at their cells again
                        if (w.rigid(k)) continue;
                        const uint32_t id[4] = {w.cell(0, k), w.cell(1, k), w.nid == 4 ? w.cell(2, k) : 0u, w.nid == 4 ? w.cell(3, k) : 0u};
                        if (!(last.comp == comp && !memcmp(last.id, id, sizeof id))) {
                            last.comp = comp;
                            memcpy(last.id, id, sizeof id);
                            auto it = table.find(last);
                            if (it == table.end()) { __atomic_store_n(&missing, 1, __ATOMIC_RELAXED); last.comp = -1; continue; }
                            last_num = it->second;
                        }
                        w.id[k] = last_num;
                    }
                }
        }
    });
    return missing ? 3 : 0;
}

int gpb_ids_scan(const uint32_t *solid, const int8_t *rigidE, const int8_t *rigidH, uint32_t *ID, int nx, int ny, int nz, int x0, int x1,
                 gpb_idcombo_t *combos, int max_combos, int *ncombos)
{
    if (!solid || !rigidE || !rigidH || !ID || !combos || !ncombos || nx < 1 || ny < 1 || nz < 1 || x0 < 0 || x1 > nx + 1 || x1 < x0) return 1;
    return ids_scan(Grid{solid, rigidE, rigidH, ID, nx, ny, nz, 0, nx, 0, nx + 1}, x0, x1, combos, max_combos, ncombos);
}

int gpb_ids_apply(const uint32_t *solid, const int8_t *rigidE, const int8_t *rigidH, uint32_t *ID, int nx, int ny, int nz, int x0, int x1,
                  const gpb_idcombo_t *combos, const uint32_t *numid, int ncombos)
{
    if (!solid || !rigidE || !rigidH || !ID || (ncombos && (!combos || !numid)) || nx < 1 || ny < 1 || nz < 1 || x0 < 0 || x1 > nx + 1 || x1 < x0) return 1;
    return ids_apply(Grid{solid, rigidE, rigidH, ID, nx, ny, nz, 0, nx, 0, nx + 1}, x0, x1, combos, numid, ncombos);
}

int gpb_ids_scan_slab(const uint32_t *solid, const int8_t *rigidE, const int8_t *rigidH, uint32_t *ID, int nx, int ny, int nz,
                      int solid_x0, int solid_nx, int id_x0, int id_nx, int x0, int x1, gpb_idcombo_t *combos, int max_combos, int *ncombos)
{
    if (!solid || !rigidE || !rigidH || !ID || !combos || !ncombos || nx < 1 || ny < 1 || nz < 1 || x0 < 0 || x1 > nx + 1 || x1 < x0) return 1;
    const Grid g{solid, rigidE, rigidH, ID, nx, ny, nz, solid_x0, solid_nx, id_x0, id_nx};
    if (solid_x0 < 0 || solid_nx < 1 || solid_x0 + solid_nx > nx || id_x0 < 0 || id_nx < 1 || id_x0 + id_nx > nx + 1 || !g.covers(x0, x1)) return 1;
    return ids_scan(g, x0, x1, combos, max_combos, ncombos);
}

int gpb_ids_apply_slab(const uint32_t *solid, const int8_t *rigidE, const int8_t *rigidH, uint32_t *ID, int nx, int ny, int nz,
                       int solid_x0, int solid_nx, int id_x0, int id_nx, int x0, int x1, const gpb_idcombo_t *combos, const uint32_t *numid, int ncombos)
{
    if (!solid || !rigidE || !rigidH || !ID || (ncombos && (!combos || !numid)) || nx < 1 || ny < 1 || nz < 1 || x0 < 0 || x1 > nx + 1 || x1 < x0) return 1;
    const Grid g{solid, rigidE, rigidH, ID, nx, ny, nz, solid_x0, solid_nx, id_x0, id_nx};
    if (solid_x0 < 0 || solid_nx < 1 || solid_x0 + solid_nx > nx || id_x0 < 0 || id_nx < 1 || id_x0 + id_nx > nx + 1 || !g.covers(x0, x1)) return 1;
    return ids_apply(g, x0, x1, combos, numid, ncombos);
}

}  // extern "C"
