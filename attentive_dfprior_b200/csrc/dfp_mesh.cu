// Mesh extraction after the Mesher's dense query: marching cubes over the occupancy lattice (include/dfprior_mesh.h;
// replaces skimage.measure.marching_cubes in src/utils/Mesher.py:463-480).  Two passes over the cells -- count, then emit at
// the offsets of the caller's prefix sum -- with the generated, face-consistent case tables of dfp_mesh_tables.h.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/dfprior_b200.h"
#include "../../include/dfprior_mesh.h"
#include "dfp_mesh_tables.h"

namespace dfp {
int fail(int code, const char* fmt, ...);          // dfp_api.cu: records the message for dfp_last_error()

struct McArgs {
    const float* vol; int nx, ny, nz; float level;
    float sx, sy, sz, ox, oy, oz;
    const int32_t* counts_in; const int64_t* offs; int32_t* counts; float* tris; int64_t* keys;
};

__device__ __forceinline__ int mc_case(const McArgs& a, int x, int y, int z) {
    int m = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float v = a.vol[((int64_t)(x + (c & 1)) * a.ny + (y + ((c >> 1) & 1))) * a.nz + (z + ((c >> 2) & 1))];
        m |= (v < a.level ? 1 : 0) << c;
    }
    return m;
}

__global__ void __launch_bounds__(256) k_mc_count(const McArgs a) {
    const int64_t cells = (int64_t)(a.nx - 1) * (a.ny - 1) * (a.nz - 1);
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (int64_t)gridDim.x * blockDim.x) {
        const int z = (int)(c % (a.nz - 1)), y = (int)((c / (a.nz - 1)) % (a.ny - 1)), x = (int)(c / ((int64_t)(a.nz - 1) * (a.ny - 1)));
        a.counts[c] = mc::TRI_COUNT[mc_case(a, x, y, z)];
    }
}

__global__ void __launch_bounds__(256) k_mc_emit(const McArgs a) {
    const int64_t cells = (int64_t)(a.nx - 1) * (a.ny - 1) * (a.nz - 1);
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (int64_t)gridDim.x * blockDim.x) {
        const int n = a.counts_in[c];
        if (n == 0) continue;
        const int z = (int)(c % (a.nz - 1)), y = (int)((c / (a.nz - 1)) % (a.ny - 1)), x = (int)(c / ((int64_t)(a.nz - 1) * (a.ny - 1)));
        const int m = mc_case(a, x, y, z);
        const int64_t base = a.offs[c];
        for (int t = 0; t < n; ++t) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int e = mc::TRI_EDGE[m][3 * t + k];
                const int ca = mc::EDGE_A[e], axis = mc::EDGE_AXIS[e];
                const int ix = x + (ca & 1), iy = y + ((ca >> 1) & 1), iz = z + ((ca >> 2) & 1);
                const int64_t ia = ((int64_t)ix * a.ny + iy) * a.nz + iz;
                const int64_t ib = ia + (axis == 0 ? (int64_t)a.ny * a.nz : axis == 1 ? a.nz : 1);
                const float va = a.vol[ia], vb = a.vol[ib];
                const float tt = __fdiv_rn(__fsub_rn(a.level, va), __fsub_rn(vb, va));
                float px = (float)ix, py = (float)iy, pz = (float)iz;
                if (axis == 0) px = __fadd_rn(px, tt); else if (axis == 1) py = __fadd_rn(py, tt); else pz = __fadd_rn(pz, tt);
                float* o = a.tris + ((base + t) * 3 + k) * 3;
                o[0] = __fadd_rn(__fmul_rn(px, a.sx), a.ox);
                o[1] = __fadd_rn(__fmul_rn(py, a.sy), a.oy);
                o[2] = __fadd_rn(__fmul_rn(pz, a.sz), a.oz);
                a.keys[(base + t) * 3 + k] = ia * 3 + axis;
            }
        }
    }
}

static int mc_check(const float* vol, int nx, int ny, int nz) {
    if (!vol) return fail(DFP_ERR_INVALID, "null volume");
    if (nx < 2 || ny < 2 || nz < 2) return fail(DFP_ERR_INVALID, "marching cubes needs at least 2 lattice points per axis (got %d x %d x %d)", nx, ny, nz);
    return 0;
}
}  // namespace dfp

using namespace dfp;

extern "C" int dfp_mc_count(const float* vol, int32_t nx, int32_t ny, int32_t nz, float level, int32_t* counts, void* stream) {
    int rc = mc_check(vol, nx, ny, nz);
    if (rc) return rc;
    if (!counts) return fail(DFP_ERR_INVALID, "null argument");
    McArgs a{};
    a.vol = vol; a.nx = nx; a.ny = ny; a.nz = nz; a.level = level; a.counts = counts;
    const int64_t cells = (int64_t)(nx - 1) * (ny - 1) * (nz - 1);
    const int blocks = (int)((cells + 255) / 256 < 148 * 16 ? (cells + 255) / 256 : 148 * 16);
    k_mc_count<<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
    if (cudaGetLastError() != cudaSuccess) return fail(DFP_ERR_CUDA, "k_mc_count launch failed");
    return 0;
}

extern "C" int dfp_mc_emit(const float* vol, int32_t nx, int32_t ny, int32_t nz, float level, const float spacing[3], const float origin[3],
                           const int32_t* counts, const int64_t* offsets, float* tris, int64_t* edge_ids, void* stream) {
    int rc = mc_check(vol, nx, ny, nz);
    if (rc) return rc;
    if (!spacing || !origin || !counts || !offsets || !tris || !edge_ids) return fail(DFP_ERR_INVALID, "null argument");
    McArgs a{};
    a.vol = vol; a.nx = nx; a.ny = ny; a.nz = nz; a.level = level;
    a.sx = spacing[0]; a.sy = spacing[1]; a.sz = spacing[2]; a.ox = origin[0]; a.oy = origin[1]; a.oz = origin[2];
    a.counts_in = counts; a.offs = offsets; a.tris = tris; a.keys = edge_ids;
    const int64_t cells = (int64_t)(nx - 1) * (ny - 1) * (nz - 1);
    const int blocks = (int)((cells + 255) / 256 < 148 * 16 ? (cells + 255) / 256 : 148 * 16);
    k_mc_emit<<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
    if (cudaGetLastError() != cudaSuccess) return fail(DFP_ERR_CUDA, "k_mc_emit launch failed");
    return 0;
}
