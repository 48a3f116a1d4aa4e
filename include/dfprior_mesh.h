/* dfprior_mesh.h -- C ABI of the mesh extraction that follows the Mesher's dense occupancy query (SURVEY.md section 8(f),
 * row 4; same library, libdfprior_b200.so).
 *
 * Replaces, in src/utils/Mesher.py:get_mesh, the call  skimage.measure.marching_cubes(volume, level, spacing)  (:463-480)
 * with two kernels over the [nx][ny][nz] float32 lattice that query_occupancy produced:
 *
 *   dfp_mc_count   one int per cell: the number of triangles its sign pattern yields
 *   (caller: exclusive prefix sum of the counts, total = number of triangles)
 *   dfp_mc_emit    the triangles: three world-space vertices each, and for every vertex the id of the lattice edge it sits on
 *                  (vertices with the same id are the same vertex: weld with a unique / inverse on the ids)
 *
 * A corner is INSIDE when its value is below `level`; triangles are oriented with their normals pointing out of the
 * inside region.  The case tables are generated (oracle/mc_tables.py): face-consistent, hence watertight, but not
 * skimage's Lewiner tables -- on cells with ambiguous faces the triangulation may differ from the reference's, and the
 * vertex / face order differs (skimage is not available in this image).
 * Vertex position on an edge from corner a (value va) to corner b (value vb), a the lower lattice index:
 *   t = (level - va) / (vb - va);  p[axis] = (index_a[axis] + t) * spacing[axis] + origin[axis]      (all float32, no FMA)
 * All pointers are device pointers; kernels run on `stream`; return 0 or a negative DfpStatus (dfp_last_error()).
 */
#ifndef DFPRIOR_MESH_H
#define DFPRIOR_MESH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

int dfp_mc_count(const float* vol, int32_t nx, int32_t ny, int32_t nz, float level, int32_t* counts /* [(nx-1)(ny-1)(nz-1)] */, void* stream);
int dfp_mc_emit(const float* vol, int32_t nx, int32_t ny, int32_t nz, float level, const float spacing[3], const float origin[3],
                const int32_t* counts, const int64_t* offsets /* exclusive prefix sum of counts */,
                float* tris /* [T][3][3] */, int64_t* edge_ids /* [T][3] */, void* stream);

#ifdef __cplusplus
}
#endif
#endif
