// scene.cu -- chunk maintenance between optimisation epochs: Morton codes for the spatial re-clustering,
// the permutation of every per-Gaussian row by the sorted order, and the chunk AABBs the frustum culling reads.
//                                  replaces the PyTorch passes of litegs/scene/point.py:22-154 and litegs/scene/cluster.py:29-46
// Not on the per-view hot path (the reference runs them every densification_interval epochs, trainer.py:103-106); they are
// HBM-bound single passes here instead of ~20 elementwise PyTorch kernels plus one indexing kernel per tensor.
#include "common.cuh"

// 3 x `bits` interleaved Morton code of each point, exactly as _gen_morton_code (point.py:38-81): fp32 normalisation
// ((p - min) / max(max - min, 1e-12)) * (2^bits - 1), truncation to integer, clamp, bit i of x/y/z to bits 3i, 3i+1, 3i+2.
__global__ void morton_kernel(const float* __restrict__ xyz, const float* __restrict__ lo, const float* __restrict__ hi, int N, int bits,
                              long long* __restrict__ codes)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float scale = (float)((1ll << bits) - 1);
    unsigned long long q[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        float denom = fmaxf(__fsub_rn(hi[a], lo[a]), 1e-12f);
        float nrm = __fmul_rn(__fdiv_rn(__fsub_rn(xyz[(size_t)a * N + i], lo[a]), denom), scale);
        long long v = (long long)nrm;                            // .long(): truncation toward zero
        v = v < 0 ? 0 : (v > (long long)scale ? (long long)scale : v);
        q[a] = (unsigned long long)v;
    }
    unsigned long long code = 0;
    for (int b = 0; b < bits; b++)
        code |= ((q[0] >> b) & 1ull) << (3 * b) | ((q[1] >> b) & 1ull) << (3 * b + 1) | ((q[2] >> b) & 1ull) << (3 * b + 2);
    codes[i] = (long long)code;
}

extern "C" int lgs_morton_codes(const float* xyz, const float* lo3, const float* hi3, int N, int bits, long long* codes, void* stream)
{
    LGS_REQUIRE(bits >= 1 && bits <= 21, "morton_codes: %d bits per axis not in 1..21", bits);
    if (N <= 0) return LGS_OK;
    morton_kernel<<<lgs_cdiv(N, 256), 256, 0, (cudaStream_t)stream>>>(xyz, lo3, hi3, N, bits, codes);
    LGS_CHECK_LAUNCH("morton_kernel");
    return LGS_OK;
}

// dst[r, j] = src[r, idx[j]] for every row r: ONE launch permutes all R rows of a [R, N] matrix (parameters, gradients and the
// two Adam moments are [rows, N] views of their storage), reading idx once per column.
__global__ void permute_rows_kernel(const float* __restrict__ src, const long long* __restrict__ idx, int R, int N, float* __restrict__ dst)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const size_t s = (size_t)idx[j];
    for (int r = blockIdx.y; r < R; r += gridDim.y) dst[(size_t)r * N + j] = src[(size_t)r * N + s];
}

extern "C" int lgs_permute_rows(const float* src, const long long* idx, int R, int N, float* dst, void* stream)
{
    LGS_REQUIRE(src != dst, "permute_rows: in-place permutation is not supported");
    if (R <= 0 || N <= 0) return LGS_OK;
    dim3 grid(lgs_cdiv(N, 256), R < 64 ? R : 64);
    permute_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, idx, R, N, dst);
    LGS_CHECK_LAUNCH("permute_rows_kernel");
    return LGS_OK;
}

// Chunk AABBs (cluster.py:29-46): per Gaussian the half extent along each world axis is sqrt(2 ln 255) * sum_a |s_a R[a][k]|
// (the S.R rows scaled: GR/transform.cu:106-125), the chunk's box is the union of [p - e, p + e].  One CTA per chunk, the
// chunk's S Gaussians reduced through shared memory.  Inputs are the RAW parameters (log scale, unnormalised quaternion).
__global__ void cluster_aabb_kernel(const float* __restrict__ xyz, const float* __restrict__ scale, const float* __restrict__ rot, int C,
                                    int S, float* __restrict__ origin, float* __restrict__ extend)
{
    __shared__ float s_lo[3][32], s_hi[3][32];
    const int c = blockIdx.x, t = threadIdx.x;
    const size_t CS = (size_t)C * S;
    float lo[3] = { 3.4028235e38f, 3.4028235e38f, 3.4028235e38f }, hi[3] = { -3.4028235e38f, -3.4028235e38f, -3.4028235e38f };
    for (int s = t; s < S; s += blockDim.x) {
        const size_t i = (size_t)c * S + s;
        float q[4] = { rot[i], rot[CS + i], rot[2 * CS + i], rot[3 * CS + i] };
        float inv = 1.0f / sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);       // F.normalize(rot, dim=0), trainer.py:106
        const float r = q[0] * inv, x = q[1] * inv, y = q[2] * inv, z = q[3] * inv;
        const float R[9] = { 1 - 2 * (y * y + z * z), 2 * (x * y + r * z), 2 * (x * z - r * y),
                             2 * (x * y - r * z), 1 - 2 * (x * x + z * z), 2 * (y * z + r * x),
                             2 * (x * z + r * y), 2 * (y * z - r * x), 1 - 2 * (x * x + y * y) };
        const float sc[3] = { expf(scale[i]), expf(scale[CS + i]), expf(scale[2 * CS + i]) };
        const float k = 3.3290429115295410f;                     // sqrt(2 ln 255)
#pragma unroll
        for (int d = 0; d < 3; d++) {
            float e = (fabsf(sc[0] * R[0 * 3 + d]) + fabsf(sc[1] * R[1 * 3 + d]) + fabsf(sc[2] * R[2 * 3 + d])) * k;
            float p = xyz[(size_t)d * CS + i];
            lo[d] = fminf(lo[d], p - e); hi[d] = fmaxf(hi[d], p + e);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; d++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
        if ((t & 31) == 0) { s_lo[d][t >> 5] = lo[d]; s_hi[d][t >> 5] = hi[d]; }
    }
    __syncthreads();
    if (t < 3) {
        const int nw = (blockDim.x + 31) >> 5;
        float a = s_lo[t][0], b = s_hi[t][0];
        for (int w = 1; w < nw; w++) { a = fminf(a, s_lo[t][w]); b = fmaxf(b, s_hi[t][w]); }
        origin[(size_t)t * C + c] = (b + a) * 0.5f;
        extend[(size_t)t * C + c] = (b - a) * 0.5f;
    }
}

extern "C" int lgs_cluster_aabb(const float* xyz, const float* scale, const float* rot, int C, int S, float* origin, float* extend,
                                void* stream)
{
    LGS_REQUIRE(S >= 1, "cluster_aabb: bad chunk size %d", S);
    if (C <= 0) return LGS_OK;
    int threads = S >= 1024 ? 1024 : ((S + 31) / 32) * 32;
    cluster_aabb_kernel<<<C, threads, 0, (cudaStream_t)stream>>>(xyz, scale, rot, C, S, origin, extend);
    LGS_CHECK_LAUNCH("cluster_aabb_kernel");
    return LGS_OK;
}
