// fused.cu -- "Level B" of the render hot path: the per-view projection chain collapsed into one kernel
// per direction.
//
// project_forward : gather visible chunk -> activate -> SH colour -> MVP -> S.R -> J -> cov2d -> inverse ->
//                   visibility + exact tile count -> 48-byte record + depth key + count.
//                   It replaces cull_compact_activate + mvp_transform_forward + createTransformMatrix_forward +
//                   jacobianRayspace + createCov2dDirectly_forward + eigh_and_inv_2x2matrix_forward +
//                   get_allocate_size + pack_forward_params (GR/compact.cu:825-893, GR/transform.cu:22-127,
//                   378-438,736-780,1364-1421, GR/binning.cu:289-385, GR/raster.cu:334-356) and never writes
//                   their ~230 B/Gaussian of intermediates (view_pos, ndc, T, J, cov2d, eig, inv, bbox) to HBM.
// project_backward: record gradient -> (recomputed chain) -> the six compacted parameter gradients; replaces
//                   unpack_gradient + inv_2x2matrix_backward + createCov2dDirectly_backward +
//                   createTransformMatrix_backward + mvp_transform_backward + activate_backward.
// emit_pairs_rec  : the (tile, splat) emission pass reading the record instead of three SoA tensors.
//
// Arithmetic is expression-for-expression the same as the op-level kernels in per_gaussian.cu / binning.cu,
// so both levels are checked against the same oracle.  One view per launch (the reference's effective
// configuration); batches of views loop on the host.
#include "common.cuh"
#include "sh.cuh"
#include "splat_geom.cuh"

struct ProjIntermediates {
    float s[3], qn[4], rn, o;       // activated scale, unit quaternion, 1/|q|, opacity
    float v[4], h[4], iw;           // view position, homogeneous clip position, 1/w
    float R[9];                     // rotation of the unit quaternion
    float VJ[6], M[6];              // V3x3.J and T.V3x3.J
    float inv[3];                   // A, B, C of the inverse 2D covariance
    float dirn[3];                  // unit view direction (for SH)
};

__device__ __forceinline__ void fused_quat_R(float r, float x, float y, float z, float* R)
{
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y + r * z);     R[2] = 2 * (x * z - r * y);
    R[3] = 2 * (x * y - r * z);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z + r * x);
    R[6] = 2 * (x * z + r * y);     R[7] = 2 * (y * z - r * x);     R[8] = 1 - 2 * (x * x + y * y);
}

// The forward chain shared by both directions.  Vm/P are the view / projection matrices (row-vector).
__device__ __forceinline__ void project_chain(const float* __restrict__ Vm, const float* __restrict__ P, const float* p,
                                              const float* s_raw, const float* q_raw, float o_raw, int H, int W,
                                              ProjIntermediates& t)
{
#pragma unroll
    for (int k = 0; k < 3; k++) t.s[k] = expf(s_raw[k]);
    t.rn = 1.0f / sqrtf(q_raw[0] * q_raw[0] + q_raw[1] * q_raw[1] + q_raw[2] * q_raw[2] + q_raw[3] * q_raw[3] + 1e-12f);
#pragma unroll
    for (int k = 0; k < 4; k++) t.qn[k] = q_raw[k] * t.rn;
    t.o = 1.0f / (1.0f + expf(-o_raw));
    float cc[3];
    lgs_camera_center(Vm, cc);
    float d0 = p[0] - cc[0], d1 = p[1] - cc[1], d2 = p[2] - cc[2];
    float dn = 1.0f / sqrtf(d0 * d0 + d1 * d1 + d2 * d2 + 1e-12f);
    t.dirn[0] = d0 * dn; t.dirn[1] = d1 * dn; t.dirn[2] = d2 * dn;
    // MVP (GR/transform.cu:398-436), w component of the world position is 1
#pragma unroll
    for (int k = 0; k < 4; k++) t.v[k] = p[0] * Vm[k] + p[1] * Vm[4 + k] + p[2] * Vm[8 + k] + 1.0f * Vm[12 + k];
#pragma unroll
    for (int k = 0; k < 4; k++) t.h[k] = t.v[0] * P[k] + t.v[1] * P[4 + k] + t.v[2] * P[8 + k] + t.v[3] * P[12 + k];
    t.iw = (fabsf(t.h[3]) > 1e-12f) ? (1.0f / t.h[3]) : 0.0f;
    // S.R (GR/transform.cu:106-125)
    fused_quat_R(t.qn[0], t.qn[1], t.qn[2], t.qn[3], t.R);
    // J (GR/transform.cu:36-50): only (0,0),(1,1),(2,0),(2,1) are non-zero
    float p00 = P[0], p11 = P[5];
    float fx = p00 * W * 0.5f, fy = p11 * H * 0.5f;
    float tx = t.v[0], ty = t.v[1], tz = t.v[2];
    float lx = tz / p00 * 1.3f, ly = tz / p11 * 1.3f;
    tx = fmaxf(fminf(tx, lx), -lx);
    ty = fmaxf(fminf(ty, ly), -ly);
    float rz = 1.0f / fmaxf(tz, 1e-2f);
    float rz2 = rz * rz;
    float J[6] = { fx * rz, 0.f, 0.f, fy * rz, -fx * tx * rz2, -fy * ty * rz2 };  // J[a*2+c], a=0..2, c=0..1
    // M = T.V3.J (GR/transform.cu:761-769), same summation order as cov_M() in per_gaussian.cu
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 3; k++) acc += Vm[a * 4 + k] * J[k * 2 + c];
            t.VJ[a * 2 + c] = acc;
        }
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 3; k++) acc += (t.R[a * 3 + k] * t.s[a]) * t.VJ[k * 2 + c];
            t.M[a * 2 + c] = acc;
        }
    float c00 = t.M[0] * t.M[0] + t.M[2] * t.M[2] + t.M[4] * t.M[4] + 0.3f;
    float c01 = t.M[0] * t.M[1] + t.M[2] * t.M[3] + t.M[4] * t.M[5];
    float c11 = t.M[1] * t.M[1] + t.M[3] * t.M[3] + t.M[5] * t.M[5] + 0.3f;
    // guarded inverse (GR/transform.cu:1379-1419)
    float det = c00 * c11 - c01 * c01;
    float det1 = (c00 - c01) * (c11 - c01) + c01 * (c00 + c11 - 2 * c01);
    det = (fabsf(det) < fabsf(1e-5f * c01 * c01)) ? det1 : det;
    det = (fabsf(det) < 1e-9f) ? 1e-9f : det;
    float dr = 1.0f / det;
    t.inv[0] = c11 * dr; t.inv[1] = -c01 * dr; t.inv[2] = c00 * dr;
}

// grid = allocated chunks (all M: chunks >= *visible_num write an invisible record), block = chunk size
template <int DEG, int TH, int TW>
__global__ void project_forward_kernel(
    const int64_t* __restrict__ chunk_ids, const int* __restrict__ visible_num, const float* __restrict__ view,
    const float* __restrict__ proj, const float* __restrict__ pos, const float* __restrict__ scale,
    const float* __restrict__ rot, const float* __restrict__ sh0, const float* __restrict__ shr,
    const float* __restrict__ opac, int C, int S, int H, int W, int gx, int gy, SplatRec* __restrict__ recs,
    unsigned* __restrict__ depth_key, unsigned* __restrict__ iota, int* __restrict__ tile_count, int* __restrict__ totals)
{
    const int a = blockIdx.x, s = threadIdx.x;
    const size_t dst = (size_t)a * S + s;
    int count = 0;
    unsigned key = 0xFFFFFFFFu;
    SplatRec r;
    r.px = r.py = 0.f; r.A = r.B = r.C = 0.f; r.o = 0.f; r.r = r.g = r.b = 0.f; r.depth = 0.f; r.pad0 = r.pad1 = 0.f;
    if (a < visible_num[0]) {
        const size_t CS = (size_t)C * S;
        const size_t src = (size_t)chunk_ids[a] * S + s;
        float p[3] = { pos[src], pos[CS + src], pos[2 * CS + src] };
        float sr_[3] = { scale[src], scale[CS + src], scale[2 * CS + src] };
        float q[4] = { rot[src], rot[CS + src], rot[2 * CS + src], rot[3 * CS + src] };
        ProjIntermediates t;
        project_chain(view, proj, p, sr_, q, opac[src], H, W, t);
        // colour (GR/compact.cu:573-653), no clamp on this path
        constexpr int K = (DEG + 1) * (DEG + 1);
        float b[16];
        lgs_sh_basis<DEG>(t.dirn[0], t.dirn[1], t.dirn[2], b);
        float col[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float acc = b[0] * sh0[c * CS + src];
#pragma unroll
            for (int k = 1; k < K; k++) acc += b[k] * shr[((size_t)(k - 1) * 3 + c) * CS + src];
            col[c] = acc + 0.5f;
        }
        const float ndcx = t.h[0] * t.iw, ndcy = t.h[1] * t.iw, ndcz = t.h[2] * t.iw;
        SplatGeom g;
        lgs_splat_setup<TH, TW>(ndcx, ndcy, t.v[2], t.inv[0], t.inv[1], t.inv[2], t.o, H, W, gx, gy, true, g);
        if (g.visible) count = lgs_process_tiles<TH, TW, false>(g, gx, 0, 0, 0, (int*)nullptr, (int*)nullptr);
        if (count > 0) key = __float_as_uint(t.v[2]);       // v.z > 0.2 here: positive floats order as unsigned
        r.px = __fsub_rn(__fmul_rn(__fmul_rn(__fadd_rn(ndcx, 1.0f), 0.5f), (float)W), 0.5f);   // GR/raster.cu:347-348
        r.py = __fsub_rn(__fmul_rn(__fmul_rn(__fadd_rn(ndcy, 1.0f), 0.5f), (float)H), 0.5f);
        r.A = t.inv[0]; r.B = t.inv[1]; r.C = t.inv[2]; r.o = t.o;
        r.r = col[0]; r.g = col[1]; r.b = col[2]; r.depth = t.v[2]; r.pad0 = ndcx; r.pad1 = ndcy;   // depth slot: view-space z (the sort key)
        (void)ndcz;
    }
    recs[dst] = r;
    depth_key[dst] = key;
    iota[dst] = (unsigned)dst;
    tile_count[dst] = count;
    // total number of (tile, splat) pairs: integer sum, order independent -> deterministic
    // totals: pair count (integer sum: order independent, deterministic) and the range of the depth keys that matter
    // (splats with pairs) so that the host can sort only the bits of that range.  totals[1] holds max(~key) so that a
    // zero fill initialises both.  One set of atomics per block.
    __shared__ int s_sum[32];
    __shared__ unsigned s_min[32], s_max[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    int wsum = __reduce_add_sync(0xffffffffu, count);
    unsigned wmin = __reduce_min_sync(0xffffffffu, count > 0 ? key : 0xFFFFFFFFu);
    unsigned wmax = __reduce_max_sync(0xffffffffu, count > 0 ? key : 0u);
    if (lane == 0) { s_sum[wid] = wsum; s_min[wid] = wmin; s_max[wid] = wmax; }
    __syncthreads();
    if (wid == 0) {
        int bs = lane < nw ? s_sum[lane] : 0;
        unsigned bmin = lane < nw ? s_min[lane] : 0xFFFFFFFFu, bmax = lane < nw ? s_max[lane] : 0u;
        bs = __reduce_add_sync(0xffffffffu, bs);
        bmin = __reduce_min_sync(0xffffffffu, bmin);
        bmax = __reduce_max_sync(0xffffffffu, bmax);
        if (lane == 0 && bs != 0) {
            atomicAdd(&totals[0], bs);
            atomicMax((unsigned*)&totals[1], ~bmin);
            atomicMax((unsigned*)&totals[2], bmax);
        }
    }
}

extern "C" int lgs_project_forward(int sh_degree, const int64_t* visible_chunk_id, const int* visible_chunks_num,
                                   const float* view_matrix, const float* proj_matrix, const float* position,
                                   const float* scale, const float* rotation, const float* sh_base, const float* sh_rest,
                                   const float* opacity, int C, int S, int A, int img_h, int img_w, int tile_h, int tile_w,
                                   float* packed_params, unsigned* depth_key, unsigned* iota, int* tile_count, int* totals,
                                   void* stream)
{
    LGS_REQUIRE(sh_degree >= 0 && sh_degree <= 3, "project_forward: sh_degree %d not in 0..3", sh_degree);
    LGS_REQUIRE(lgs_tile_ok(tile_h, tile_w), "project_forward: tile %dx%d not one of 8x16, 12x16, 16x16, 8x8", tile_h, tile_w);
    LGS_REQUIRE(S >= 32 && S <= 1024 && S % 32 == 0, "project_forward: chunk size %d must be a multiple of 32 in 32..1024", S);
    cudaStream_t st = (cudaStream_t)stream;
    LGS_CUDA(cudaMemsetAsync(totals, 0, 3 * sizeof(int), st));
    if (A == 0) return LGS_OK;
    int gx = (img_w + tile_w - 1) / tile_w, gy = (img_h + tile_h - 1) / tile_h;
#define PF(D) project_forward_kernel<D, TH, TW><<<A, S, 0, st>>>(visible_chunk_id, visible_chunks_num, view_matrix, proj_matrix, position, \
        scale, rotation, sh_base, sh_rest, opacity, C, S, img_h, img_w, gx, gy, (SplatRec*)packed_params, depth_key, iota, tile_count, totals)
    LGS_DISPATCH_TILE(tile_h, tile_w,
        switch (sh_degree) { case 0: PF(0); break; case 1: PF(1); break; case 2: PF(2); break; default: PF(3); })
#undef PF
    LGS_CHECK_LAUNCH("project_forward_kernel");
    return LGS_OK;
}

// (tile+1, splat) emission in depth order from the packed record.    replaces GR/binning.cu:33-110
// One splat per lane.  A warp's 32 splats own one contiguous run [wbase, wend) of the pair list (the offsets
// are a scan in the same order), so the lanes stage their pairs in a per-warp shared-memory window and the
// warp then copies the window out with fully coalesced stores; without the staging every lane streams its own
// run and the kernel sits on the LSU queue (ncu: lg_throttle was the top stall).  Runs longer than the window
// are handled by walking again per window (rare: near-camera splats).
constexpr int LGS_EMIT_WARPS = 8;
#ifndef LGS_EMIT_WINDOW
#define LGS_EMIT_WINDOW 512
#endif
template <int TH, int TW, typename KeyT>
__global__ void __launch_bounds__(LGS_EMIT_WARPS * 32) emit_pairs_rec_kernel(const SplatRec* __restrict__ recs, const int* __restrict__ offset,
                                                             const unsigned* __restrict__ order, int n, int cap, int H, int W,
                                                             int gx, int gy, KeyT* __restrict__ keys, int* __restrict__ vals,
                                                             const int* __restrict__ n_dev, int* __restrict__ valid_pairs)
{
    __shared__ KeyT s_keys[LGS_EMIT_WARPS][LGS_EMIT_WINDOW];
    __shared__ int s_vals[LGS_EMIT_WARPS][LGS_EMIT_WINDOW];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (n_dev != nullptr) n = min(n, max(*n_dev, 0));         // GPU-driven sizing: n is the capacity, *n_dev the live count
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j - lane >= n) return;                               // whole warp past the end
    int off = 0, asz = 0, i = 0;
    if (j < n) {
        off = j == 0 ? 0 : offset[j - 1];
        asz = offset[j] - off;
        // the first run that crosses the capacity: everything from here on is dropped, so the list is valid up to `off`
        // exactly (offsets are monotone: this thread is unique) -- the sort and the tile ranges must not look further
        if (valid_pairs != nullptr && asz > 0 && off <= cap && off + asz > cap) *valid_pairs = off;
        if (asz <= 0 || off + asz > cap) asz = 0;
        i = (int)order[j];
    }
    SplatGeom g;
    g.visible = false;
    if (asz > 0) {
        const SplatRec r = recs[i];
        lgs_splat_setup<TH, TW>(r.pad0, r.pad1, 1.0f, r.A, r.B, r.C, r.o, H, W, gx, gy, false, g);
    }
    const bool live = asz > 0 && g.visible;
    const int wbase = __reduce_min_sync(0xffffffffu, live ? off : 0x7fffffff);
    const int wend = __reduce_max_sync(0xffffffffu, live ? off + asz : 0);
    KeyT* sk = s_keys[wid];
    int* sv = s_vals[wid];
    for (int lo = wbase; lo < wend; lo += LGS_EMIT_WINDOW) {
        if (live && off < lo + LGS_EMIT_WINDOW && off + asz > lo)
            lgs_process_tiles<TH, TW, true, KeyT>(g, gx, i, off, LGS_EMIT_WINDOW, sk, sv, lo);
        __syncwarp();
        const int m = min(LGS_EMIT_WINDOW, wend - lo);
        for (int k = lane; k < m; k += 32) { keys[lo + k] = sk[k]; vals[lo + k] = sv[k]; }
        __syncwarp();
    }
}

extern "C" int lgs_emit_pairs(const float* packed_params, const int* offset, const unsigned* order, int n, int cap, int img_h,
                              int img_w, int tile_h, int tile_w, int* keys, int* vals, void* stream)
{
    LGS_REQUIRE(lgs_tile_ok(tile_h, tile_w), "emit_pairs: tile %dx%d not one of 8x16, 12x16, 16x16, 8x8", tile_h, tile_w);
    if (n <= 0 || cap <= 0) return LGS_OK;
    int gx = (img_w + tile_w - 1) / tile_w, gy = (img_h + tile_h - 1) / tile_h;
    cudaStream_t st = (cudaStream_t)stream;
    LGS_DISPATCH_TILE(tile_h, tile_w,
        emit_pairs_rec_kernel<TH, TW, int><<<lgs_cdiv(n, LGS_EMIT_WARPS * 32), LGS_EMIT_WARPS * 32, 0, st>>>((const SplatRec*)packed_params, offset, order, n, cap, img_h,
                                                                           img_w, gx, gy, keys, vals, nullptr, nullptr);)
    LGS_CHECK_LAUNCH("emit_pairs_rec_kernel");
    return LGS_OK;
}

// same with 16-bit tile keys (requires tiles + 1 < 65536)
extern "C" int lgs_emit_pairs_u16(const float* packed_params, const int* offset, const unsigned* order, int n, int cap, int img_h,
                                  int img_w, int tile_h, int tile_w, unsigned short* keys, int* vals, void* stream)
{
    LGS_REQUIRE(lgs_tile_ok(tile_h, tile_w), "emit_pairs_u16: tile %dx%d not one of 8x16, 12x16, 16x16, 8x8", tile_h, tile_w);
    if (n <= 0 || cap <= 0) return LGS_OK;
    int gx = (img_w + tile_w - 1) / tile_w, gy = (img_h + tile_h - 1) / tile_h;
    LGS_REQUIRE(gx * gy + 1 < 65536, "emit_pairs_u16: %d tiles do not fit 16-bit keys", gx * gy);
    cudaStream_t st = (cudaStream_t)stream;
    LGS_DISPATCH_TILE(tile_h, tile_w,
        emit_pairs_rec_kernel<TH, TW, unsigned short><<<lgs_cdiv(n, LGS_EMIT_WARPS * 32), LGS_EMIT_WARPS * 32, 0, st>>>((const SplatRec*)packed_params, offset, order, n, cap,
                                                                                      img_h, img_w, gx, gy, keys, vals, nullptr, nullptr);)
    LGS_CHECK_LAUNCH("emit_pairs_rec_kernel<u16>");
    return LGS_OK;
}

// GPU-driven forms: n_capacity bounds the launch, *n_dev is the live splat count; runs that would cross `cap` are dropped (and
// flagged by lgs_view_params) and *valid_pairs (nullable; normally &params[1]) is lowered to the length of the list that WAS
// written, so that nothing downstream reads an unwritten slot.  key_bits = 16 or 32.
extern "C" int lgs_emit_pairs_dev(const float* packed_params, const int* offset, const unsigned* order, int n_capacity, const int* n_dev,
                                  int cap, int img_h, int img_w, int tile_h, int tile_w, int key_bits, void* keys, int* vals,
                                  int* valid_pairs, void* stream)
{
    LGS_REQUIRE(lgs_tile_ok(tile_h, tile_w), "emit_pairs_dev: tile %dx%d not one of 8x16, 12x16, 16x16, 8x8", tile_h, tile_w);
    LGS_REQUIRE(n_dev != nullptr && (key_bits == 16 || key_bits == 32), "emit_pairs_dev: bad arguments");
    if (n_capacity <= 0 || cap <= 0) return LGS_OK;
    int gx = (img_w + tile_w - 1) / tile_w, gy = (img_h + tile_h - 1) / tile_h;
    LGS_REQUIRE(key_bits == 32 || gx * gy + 1 < 65536, "emit_pairs_dev: %d tiles do not fit 16-bit keys", gx * gy);
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = lgs_cdiv(n_capacity, LGS_EMIT_WARPS * 32);
    if (key_bits == 16) {
        LGS_DISPATCH_TILE(tile_h, tile_w,
            emit_pairs_rec_kernel<TH, TW, unsigned short><<<grid, LGS_EMIT_WARPS * 32, 0, st>>>((const SplatRec*)packed_params, offset, order, n_capacity,
                                                                                          cap, img_h, img_w, gx, gy, (unsigned short*)keys, vals, n_dev, valid_pairs);)
    } else {
        LGS_DISPATCH_TILE(tile_h, tile_w,
            emit_pairs_rec_kernel<TH, TW, int><<<grid, LGS_EMIT_WARPS * 32, 0, st>>>((const SplatRec*)packed_params, offset, order, n_capacity, cap,
                                                                               img_h, img_w, gx, gy, (int*)keys, vals, n_dev, valid_pairs);)
    }
    LGS_CHECK_LAUNCH("emit_pairs_rec_kernel(dev)");
    return LGS_OK;
}

// Derived launch parameters of one view, computed ON THE DEVICE from the counters project_forward leaves behind, so that the rest
// of the view can be enqueued without reading anything back:
//   params[0] = live splat slots (visible chunks * S)      params[1] = pairs, clamped to the pair capacity
//   params[2] = depth-key bias (smallest key of a splat that owns pairs)
//   params[3] = bits of the depth-key range                 params[4] = flags: 1 = pairs exceed the capacity (list truncated),
//   params[5] = pairs (unclamped)                                        2 = depth-key range needs more bits than planned
//   params[6] = visible chunks                              params[7] = planned depth bits
// counters = i32[4] as written by lgs_frustum_culling_aabb ([0]) and lgs_project_forward ([1..3]).
// sticky (i32[4], nullable) accumulates over views: [0] |= flags, [1] = max pairs, [2] = max depth bits, [3] += 1 -- one read-back
// per BATCH of views tells whether any of them overflowed and how large the next workspace has to be.
__global__ void view_params_kernel(const int* __restrict__ counters, int S, int pair_capacity, int planned_bits, int* __restrict__ params,
                                   int* __restrict__ sticky)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int nvis = counters[0], D = counters[1];
    const unsigned kmin = ~(unsigned)counters[2], kmax = (unsigned)counters[3];
    int bits = 1;
    if (D > 0 && kmax >= kmin) { unsigned r = kmax - kmin; bits = (r == 0) ? 1 : (32 - __clz(r)); }
    int flags = 0;
    if (D > pair_capacity) flags |= 1;
    if (bits > planned_bits) flags |= 2;
    params[0] = nvis * S; params[1] = min(D, pair_capacity); params[2] = (int)kmin; params[3] = bits;
    params[4] = flags; params[5] = D; params[6] = nvis; params[7] = planned_bits;
    if (sticky != nullptr) {
        atomicOr(&sticky[0], flags); atomicMax(&sticky[1], D); atomicMax(&sticky[2], bits); atomicAdd(&sticky[3], 1);
    }
}

extern "C" int lgs_view_params(const int* counters, int S, int pair_capacity, int planned_depth_bits, int* params, int* sticky,
                               void* stream)
{
    LGS_REQUIRE(counters != nullptr && params != nullptr && planned_depth_bits >= 1 && planned_depth_bits <= 32, "view_params: bad arguments");
    view_params_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(counters, S, pair_capacity, planned_depth_bits, params, sticky);
    LGS_CHECK_LAUNCH("view_params_kernel");
    return LGS_OK;
}

__device__ __forceinline__ float nan_to_num0(float x)
{
    if (x != x) return 0.0f;                               // torch.nan_to_num_(0) of wrapper.py:591
    if (isinf(x)) return x > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
    return x;
}

template <int DEG>
__global__ void project_backward_kernel(
    const int64_t* __restrict__ chunk_ids, const int* __restrict__ visible_num, const float* __restrict__ view,
    const float* __restrict__ proj, const float* __restrict__ pos, const float* __restrict__ scale,
    const float* __restrict__ rot, const float* __restrict__ opac, int C, int S, int A, int rest_dim, int H, int W,
    int true_sigmoid, int accumulate, const float* __restrict__ grad /*[A*S,12]*/, const float* __restrict__ inv_scaler,
    float* __restrict__ g_pos, float* __restrict__ g_scale, float* __restrict__ g_rot, float* __restrict__ g_sh0,
    float* __restrict__ g_shr, float* __restrict__ g_opac, float* __restrict__ touched)
{
    const int a = blockIdx.x, s = threadIdx.x;
    if (a >= visible_num[0]) return;
    const size_t CS = (size_t)C * S, AS = (size_t)A * S;
    const size_t dst = (size_t)a * S + s, src = (size_t)chunk_ids[a] * S + s;
    if (touched != nullptr && s == 0) touched[chunk_ids[a]] = 1.0f;     // chunk mark for the fused optimizer step
    constexpr int K = (DEG + 1) * (DEG + 1);
    const float sc = inv_scaler ? inv_scaler[0] : 1.0f;
    const float4* g4 = reinterpret_cast<const float4*>(grad + dst * LGS_GRAD_FLOATS);
    const float4 ga = g4[0], gb = g4[1], gc = g4[2];
    const bool any = (ga.x != 0.f) | (ga.y != 0.f) | (ga.z != 0.f) | (ga.w != 0.f) | (gb.x != 0.f) | (gb.y != 0.f) |
                     (gb.z != 0.f) | (gb.w != 0.f) | (gc.x != 0.f);
    float o_pos[3] = { 0.f, 0.f, 0.f }, o_sc[3] = { 0.f, 0.f, 0.f }, o_q[4] = { 0.f, 0.f, 0.f, 0.f }, o_op = 0.f;
    float shb[16], dcol3[3] = { 0.f, 0.f, 0.f };       // SH basis and colour gradient: d sh[k][c] = shb[k] * dcol3[c]
#pragma unroll
    for (int k = 0; k < 16; k++) shb[k] = 0.f;
    if (any) {
        float p[3] = { pos[src], pos[CS + src], pos[2 * CS + src] };
        float sr_[3] = { scale[src], scale[CS + src], scale[2 * CS + src] };
        float q[4] = { rot[src], rot[CS + src], rot[2 * CS + src], rot[3 * CS + src] };
        const float o_raw = opac[src];
        ProjIntermediates t;
        project_chain(view, proj, p, sr_, q, o_raw, H, W, t);
        // raw moments -> record gradient (GR/raster.cu:826-841), then unpack (GR/raster.cu:870-884)
        LgsRasterGrad rgd;
        lgs_finish_raster_grad(ga, gb, gc, t.inv[0], t.inv[1], t.inv[2], t.o, rgd);
        const float d_ndcx = rgd.dmx * 0.5f * W * sc, d_ndcy = rgd.dmy * 0.5f * H * sc;
        const float dA = rgd.dA * sc, dBh = rgd.dB * 0.5f * sc, dC = rgd.dC * sc;
        const float dcol[3] = { gb.y * sc, gb.z * sc, gb.w * sc };
        const float d_o = rgd.dop * sc;
        // inverse backward: dCov = -(inv . dInv . inv) (GR/transform.cu:1446-1450), NaN -> 0
        const float iA = t.inv[0], iB = t.inv[1], iC = t.inv[2];
        float t00 = iA * dA + iB * dBh, t01 = iA * dBh + iB * dC, t10 = iB * dA + iC * dBh, t11 = iB * dBh + iC * dC;
        float G[4];
        G[0] = nan_to_num0(-(t00 * iA + t01 * iB)); G[1] = nan_to_num0(-(t00 * iB + t01 * iC));
        G[2] = nan_to_num0(-(t10 * iA + t11 * iB)); G[3] = nan_to_num0(-(t10 * iB + t11 * iC));
        // cov2d backward: dT = 2 M G (VJ)^T (GR/transform.cu:861-880)
        float dT[9];
#pragma unroll
        for (int r = 0; r < 3; r++) {
            float dM0 = 2.f * (t.M[r * 2] * G[0] + t.M[r * 2 + 1] * G[2]);
            float dM1 = 2.f * (t.M[r * 2] * G[1] + t.M[r * 2 + 1] * G[3]);
#pragma unroll
            for (int k = 0; k < 3; k++) dT[r * 3 + k] = dM0 * t.VJ[k * 2] + dM1 * t.VJ[k * 2 + 1];
        }
        // transform backward (GR/transform.cu:185-226) on the ACTIVATED scale / unit quaternion
        float ds[3];
#pragma unroll
        for (int r = 0; r < 3; r++) {
            ds[r] = t.R[r * 3] * dT[r * 3] + t.R[r * 3 + 1] * dT[r * 3 + 1] + t.R[r * 3 + 2] * dT[r * 3 + 2];
            dT[r * 3] *= t.s[r]; dT[r * 3 + 1] *= t.s[r]; dT[r * 3 + 2] *= t.s[r];
        }
        const float r_ = t.qn[0], x = t.qn[1], y = t.qn[2], z = t.qn[3];
        float dq[4];
        dq[0] = 2 * z * (dT[1] - dT[3]) + 2 * y * (dT[6] - dT[2]) + 2 * x * (dT[5] - dT[7]);
        dq[1] = 2 * y * (dT[3] + dT[1]) + 2 * z * (dT[6] + dT[2]) + 2 * r_ * (dT[5] - dT[7]) - 4 * x * (dT[8] + dT[4]);
        dq[2] = 2 * x * (dT[3] + dT[1]) + 2 * r_ * (dT[6] - dT[2]) + 2 * z * (dT[5] + dT[7]) - 4 * y * (dT[8] + dT[0]);
        dq[3] = 2 * r_ * (dT[1] - dT[3]) + 2 * x * (dT[6] + dT[2]) + 2 * y * (dT[5] + dT[7]) - 4 * z * (dT[4] + dT[0]);
        // activation chain (GR/compact.cu:925-952)
#pragma unroll
        for (int k = 0; k < 3; k++) o_sc[k] = t.s[k] * ds[k];
        const float dot = dq[0] * t.qn[0] + dq[1] * t.qn[1] + dq[2] * t.qn[2] + dq[3] * t.qn[3];
#pragma unroll
        for (int k = 0; k < 4; k++) o_q[k] = t.rn * (dq[k] - dot * t.qn[k]);
        const float sig = 1.0f - 1.0f / (1.0f + expf(o_raw));
        o_op = d_o * (true_sigmoid ? sig * (1.0f - sig) : sig);
        // MVP backward (GR/transform.cu:517-558) with d_ndc.z = d_ndc.w = 0 and no view-space gradient
        const float* P = proj; const float* Vm = view;
        const float n0 = t.h[0] * t.iw, n1 = t.h[1] * t.iw;
        float dh[4] = { d_ndcx * t.iw, d_ndcy * t.iw, 0.f * t.iw, -(d_ndcx * n0 + d_ndcy * n1 + 0.f * (t.h[2] * t.iw)) * t.iw };
        float dv[4];
#pragma unroll
        for (int k = 0; k < 4; k++) dv[k] = dh[0] * P[k * 4] + dh[1] * P[k * 4 + 1] + dh[2] * P[k * 4 + 2] + dh[3] * P[k * 4 + 3] + 0.f;
#pragma unroll
        for (int k = 0; k < 3; k++) o_pos[k] = dv[0] * Vm[k * 4] + dv[1] * Vm[k * 4 + 1] + dv[2] * Vm[k * 4 + 2] + dv[3] * Vm[k * 4 + 3];
        // SH coefficients (GR/compact.cu:655-823); the direction is treated as constant
        lgs_sh_basis<DEG>(t.dirn[0], t.dirn[1], t.dirn[2], shb);
        dcol3[0] = dcol[0]; dcol3[1] = dcol[1]; dcol3[2] = dcol[2];
    }
    if (accumulate) {
        // dense accumulation: outputs are the full [..,C,S] gradient tensors, this view's contribution is added at
        // the SOURCE chunk (each Gaussian is owned by exactly one thread of one launch: no atomics needed);
        // Gaussians that received no gradient are not touched at all.
        // The read-modify-writes are issued in batches (all loads of a batch first, then the stores) so that 11-16
        // independent L2 round trips are in flight per thread instead of one dependent load->add->store chain each.
        if (any) {
            float old[11];
#pragma unroll
            for (int k = 0; k < 3; k++) old[k] = g_pos[k * CS + src];
#pragma unroll
            for (int k = 0; k < 3; k++) old[3 + k] = g_scale[k * CS + src];
#pragma unroll
            for (int k = 0; k < 4; k++) old[6 + k] = g_rot[k * CS + src];
            old[10] = g_opac[src];
#pragma unroll
            for (int k = 0; k < 3; k++) g_pos[k * CS + src] = old[k] + o_pos[k];
#pragma unroll
            for (int k = 0; k < 3; k++) g_scale[k * CS + src] = old[3 + k] + o_sc[k];
#pragma unroll
            for (int k = 0; k < 4; k++) g_rot[k * CS + src] = old[6 + k] + o_q[k];
            g_opac[src] = old[10] + o_op;
#pragma unroll
            for (int c = 0; c < 3; c++) {
                float osh[K];
                osh[0] = g_sh0[c * CS + src];
#pragma unroll
                for (int k = 1; k < K; k++) osh[k] = g_shr[((size_t)(k - 1) * 3 + c) * CS + src];
                g_sh0[c * CS + src] = fmaf(shb[0], dcol3[c], osh[0]);
#pragma unroll
                for (int k = 1; k < K; k++) g_shr[((size_t)(k - 1) * 3 + c) * CS + src] = fmaf(shb[k], dcol3[c], osh[k]);
            }
        }
        return;
    }
#pragma unroll
    for (int k = 0; k < 3; k++) g_pos[k * AS + dst] = o_pos[k];
#pragma unroll
    for (int k = 0; k < 3; k++) g_scale[k * AS + dst] = o_sc[k];
#pragma unroll
    for (int k = 0; k < 4; k++) g_rot[k * AS + dst] = o_q[k];
    g_opac[dst] = o_op;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        g_sh0[c * AS + dst] = shb[0] * dcol3[c];
#pragma unroll
        for (int k = 1; k < K; k++) g_shr[((size_t)(k - 1) * 3 + c) * AS + dst] = shb[k] * dcol3[c];
    }
    (void)rest_dim;
}

// mode 0: outputs are compacted [..,A,S] and assigned; mode 1: same, cleared first (rows of chunks >= *visible_num and
// sh_rest rows above the active degree must read as zero); mode 2: outputs are the DENSE [..,C,S] gradient tensors and
// this view's gradients are accumulated into them (the multi-view / data-parallel path: no compacted round trip).
extern "C" int lgs_project_backward(int sh_degree, const int64_t* visible_chunk_id, const int* visible_chunks_num,
                                    const float* view_matrix, const float* proj_matrix, const float* position,
                                    const float* scale, const float* rotation, const float* opacity, int C, int S, int A,
                                    int rest_dim, int img_h, int img_w, int true_sigmoid_grad, const float* packed_grad,
                                    const float* grad_inv_scaler, int zero_outputs, float* g_position, float* g_scale,
                                    float* g_rotation, float* g_sh_base, float* g_sh_rest, float* g_opacity, float* touched,
                                    void* stream)
{
    LGS_REQUIRE(sh_degree >= 0 && sh_degree <= 3, "project_backward: sh_degree %d not in 0..3", sh_degree);
    LGS_REQUIRE(rest_dim >= (sh_degree + 1) * (sh_degree + 1) - 1, "project_backward: sh_rest has %d rows, degree %d needs %d", rest_dim,
                sh_degree, (sh_degree + 1) * (sh_degree + 1) - 1);
    LGS_REQUIRE(S >= 1 && S <= 1024, "project_backward: chunk size %d unsupported", S);
    if (A == 0) return LGS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    size_t AS = (size_t)A * S;
    const int accumulate = (zero_outputs == 2) ? 1 : 0;
    if (zero_outputs == 1) {
        LGS_CUDA(cudaMemsetAsync(g_position, 0, sizeof(float) * 3 * AS, st));
        LGS_CUDA(cudaMemsetAsync(g_scale, 0, sizeof(float) * 3 * AS, st));
        LGS_CUDA(cudaMemsetAsync(g_rotation, 0, sizeof(float) * 4 * AS, st));
        LGS_CUDA(cudaMemsetAsync(g_sh_base, 0, sizeof(float) * 3 * AS, st));
        LGS_CUDA(cudaMemsetAsync(g_sh_rest, 0, sizeof(float) * (size_t)rest_dim * 3 * AS, st));
        LGS_CUDA(cudaMemsetAsync(g_opacity, 0, sizeof(float) * AS, st));
    }
#define PB(D) project_backward_kernel<D><<<A, S, 0, st>>>(visible_chunk_id, visible_chunks_num, view_matrix, proj_matrix, position, scale, \
        rotation, opacity, C, S, A, rest_dim, img_h, img_w, true_sigmoid_grad, accumulate, packed_grad, grad_inv_scaler, g_position, g_scale, \
        g_rotation, g_sh_base, g_sh_rest, g_opacity, touched)
    switch (sh_degree) { case 0: PB(0); break; case 1: PB(1); break; case 2: PB(2); break; default: PB(3); }
#undef PB
    LGS_CHECK_LAUNCH("project_backward_kernel");
    return LGS_OK;
}
