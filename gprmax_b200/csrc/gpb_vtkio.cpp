// gpb_vtkio.cpp -- host-side re-layout for the streaming VTK writers.
//
// The reference's output files are in ParaView's order -- x fastest, z slowest, vector components interleaved -- while every
// array of the model is z fastest.  The reference does the re-layout with a serial Cython triple loop whose inner index runs
// along the slowest memory axis (define_normal_geometry, gprMax/geometry_outputs_ext.pyx:81-110: one cache line per element)
// and with `np.stack((Ex, Ey, Ez)).reshape(-1, order='F')` for snapshots (gprMax/snapshots.py:128-130, 223-228), both over
// the whole volume at once.  Here it is ONE blocked transposition that runs on all host cores and produces any z range of
// the output, so a writer can stream a file of any size through a buffer of a few k planes
// (gprmax_b200/vtk_writers.py).
#include "../../include/gprmax_b200.h"

#include <algorithm>
#include <cstring>
#include <thread>
#include <unistd.h>
#include <vector>

namespace {

constexpr int TILE = 32;   // 32 x 32 elements: 128-byte lines on both sides for 4-byte elements, tile <= 24 KB for 3 x 8 bytes

template <typename T>
void transpose_items(const T *const *src, int ncomp, const int64_t *stride, const int32_t *start, const int32_t *count, const int32_t *step, T *out,
                     long long item0, long long item1)
{
    // work item = (k tile, j): all i of one output row range.  out[((k * n1 + j) * n0 + i) * ncomp + c]
    const int n0 = count[0], n1 = count[1], n2 = count[2];
    T tile[3][TILE][TILE];
    for (long long item = item0; item < item1; ++item) {
        const int kt = (int)(item / n1), j = (int)(item % n1);
        const int k0 = kt * TILE, kn = std::min(TILE, n2 - k0);
        const int64_t joff = (int64_t)(start[1] + (int64_t)j * step[1]) * stride[1];
        for (int i0 = 0; i0 < n0; i0 += TILE) {
            const int in = std::min(TILE, n0 - i0);
            for (int c0 = 0; c0 < ncomp; c0 += 3) {
                const int cn = std::min(3, ncomp - c0);
                for (int c = 0; c < cn; ++c) {
                    const T *s = src[c0 + c] + joff + (int64_t)(start[2] + (int64_t)k0 * step[2]) * stride[2];
                    const int64_t kstride = (int64_t)step[2] * stride[2];
                    for (int ii = 0; ii < in; ++ii) {
                        const T *row = s + (int64_t)(start[0] + (int64_t)(i0 + ii) * step[0]) * stride[0];
                        for (int kk = 0; kk < kn; ++kk) tile[c][kk][ii] = row[kk * kstride];
                    }
                }
                for (int kk = 0; kk < kn; ++kk) {
                    T *o = out + (((int64_t)(k0 + kk) * n1 + j) * n0 + i0) * ncomp + c0;
                    if (ncomp == 1) {
                        memcpy(o, tile[0][kk], sizeof(T) * in);
                    } else {
                        for (int ii = 0; ii < in; ++ii)
                            for (int c = 0; c < cn; ++c) o[(int64_t)ii * ncomp + c] = tile[c][kk][ii];
                    }
                }
            }
        }
    }
}

int host_threads()
{
    const long n = sysconf(_SC_NPROCESSORS_CONF);
    return (int)std::max(1l, std::min(64l, n));
}

template <typename T>
void transpose_all(const void *const *src, int ncomp, const int64_t *stride, const int32_t *start, const int32_t *count, const int32_t *step, void *out)
{
    const long long items = (long long)((count[2] + TILE - 1) / TILE) * count[1];
    const long long elems = (long long)count[0] * count[1] * count[2];
    const int nt = (int)std::max(1ll, std::min<long long>(elems < (1 << 16) ? 1 : host_threads(), items));
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) {
        const long long a = items * t / nt, b = items * (t + 1) / nt;
        auto work = [=] { transpose_items<T>((const T *const *)src, ncomp, stride, start, count, step, (T *)out, a, b); };
        if (nt == 1) { work(); break; }
        th.emplace_back([=] {
            cpu_set_t all;   // the caller is usually pinned to one core by the reference's OMP_PROC_BIND (see gpb_idbuild.cpp)
            CPU_ZERO(&all);
            for (int c = 0; c < CPU_SETSIZE; ++c) CPU_SET(c, &all);
            pthread_setaffinity_np(pthread_self(), sizeof all, &all);
            work();
        });
    }
    for (auto &t : th) t.join();
}

}  // namespace

extern "C" int gpb_vtk_transpose(const void *const *src, int ncomp, int elem_bytes, const int64_t stride[3], const int32_t start[3],
                                 const int32_t count[3], const int32_t step[3], void *out)
{
    if (!src || !stride || !start || !count || !step || !out || ncomp < 1) return 1;
    for (int c = 0; c < ncomp; ++c)
        if (!src[c]) return 1;
    for (int a = 0; a < 3; ++a)
        if (count[a] < 0 || step[a] < 1 || start[a] < 0) return 1;
    if (count[0] == 0 || count[1] == 0 || count[2] == 0) return 0;
    switch (elem_bytes) {
    case 1: transpose_all<uint8_t>(src, ncomp, stride, start, count, step, out); return 0;
    case 2: transpose_all<uint16_t>(src, ncomp, stride, start, count, step, out); return 0;
    case 4: transpose_all<uint32_t>(src, ncomp, stride, start, count, step, out); return 0;
    case 8: transpose_all<uint64_t>(src, ncomp, stride, start, count, step, out); return 0;
    default: return 1;
    }
}
