// ssim.cu -- fused SSIM / L1+SSIM loss maps and their backward: the step either side of the rasterizer in a
// training iteration (consumes the rendered image, produces dL/dimage).
// replaces fused_ssim/ssim.cu:64-274, 277-437 (SSIM), 528-712, 719-850 (L1 + SSIM loss) and their hosts :444-524, 855-942.
//
// Same maths as the reference (11-tap sigma-1.5 Gaussian, zero padding, separable, per channel; the partial derivatives
// dm/dmu1, dm/dsigma1^2, dm/dsigma12 are saved by the forward and filtered again by the backward), different schedule:
// a CTA owns a 64 x 32 output tile.  The (64+10) x (32+10) input window of both images is staged in shared memory once;
// the horizontal pass gives every lane a strip of 8 adjacent outputs of ONE row (18 loads feed 8 x 5 x 11 FMAs, the
// products x^2, y^2, xy are formed once per pixel instead of once per tap; lanes of a warp walk consecutive rows with an
// odd row stride, so the strip loads are bank-conflict free); the vertical pass gives every thread a strip of 8 outputs
// of one column (18 x 5 loads, lanes on consecutive columns).  That is ~30 shared-memory accesses per output instead of
// the ~130 of a thread-per-pixel schedule, which is what bounds this filter on B200 (the data is small: 20 B/pixel).
// The loss map itself is optional: the training path only needs its mean, so the kernel can emit one partial sum per
// CTA instead (summed on the host side in a fixed order -> deterministic).
#include "common.cuh"

namespace {

constexpr int TX = 64, TY = 32, HALO = 5;
constexpr int IX = TX + 2 * HALO, IY = TY + 2 * HALO;       // 74 x 42 input window
constexpr int ISTRIDE = IX + 1;                              // 75: odd, lanes on consecutive rows hit distinct banks
constexpr int STRIP = 8;
constexpr int THREADS = 256;

__device__ __forceinline__ float gauss(int k)
{
    // fused_ssim/ssim.cu:12-24 (normalised exp(-(k-5)^2 / 4.5), rounded to fp32)
    constexpr float G[11] = { 0.001028380123898387f, 0.0075987582094967365f, 0.036000773310661316f, 0.10936068743467331f,
                              0.21300552785396576f,  0.26601171493530273f,   0.21300552785396576f,  0.10936068743467331f,
                              0.036000773310661316f, 0.0075987582094967365f, 0.001028380123898387f };
    return G[k];
}

// out[j] = sum_k g[k] * in[j + k]   for STRIP outputs from STRIP + 10 inputs
__device__ __forceinline__ void conv_strip(const float (&in)[STRIP + 10], float (&out)[STRIP])
{
#pragma unroll
    for (int j = 0; j < STRIP; j++) {
        float acc = gauss(0) * in[j];
#pragma unroll
        for (int k = 1; k < 11; k++) acc = fmaf(gauss(k), in[j + k], acc);
        out[j] = acc;
    }
}

__device__ __forceinline__ void load_window(const float* __restrict__ plane, int H, int W, int x0, int y0, float* __restrict__ s)
{
    for (int i = threadIdx.x; i < IY * IX; i += THREADS) {
        int ly = i / IX, lx = i - ly * IX;
        int gy = y0 - HALO + ly, gx = x0 - HALO + lx;
        float v = 0.0f;
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) v = plane[(size_t)gy * W + gx];
        s[ly * ISTRIDE + lx] = v;
    }
}

template <int NCH>
struct HBuf { float v[NCH][IY][TX]; };

// L1: loss map w (1 - ssim) + (1 - w) |x - y| instead of the ssim map.   TRAIN: write the three partials.
template <bool L1, bool TRAIN>
__global__ void __launch_bounds__(THREADS) ssim_forward_kernel(const float* __restrict__ img1, const float* __restrict__ img2, int H, int W,
                                                               float C1, float C2, float ssim_weight, float* __restrict__ map,
                                                               float* __restrict__ dm_dmu1, float* __restrict__ dm_dsigma1_sq,
                                                               float* __restrict__ dm_dsigma12, float* __restrict__ block_sums)
{
    extern __shared__ float smem[];
    float* sX = smem;
    float* sY = smem + IY * ISTRIDE;
    HBuf<5>& hb = *reinterpret_cast<HBuf<5>*>(smem + 2 * IY * ISTRIDE);
    const size_t plane = (size_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    load_window(img1 + plane, H, W, x0, y0, sX);
    load_window(img2 + plane, H, W, x0, y0, sY);
    __syncthreads();
    // horizontal pass: warp w owns the strip of columns [8w, 8w+8), lanes walk the rows
    for (int row = lane; row < IY; row += 32) {
        float a[STRIP + 10], b[STRIP + 10], p[STRIP + 10], o[STRIP];
#pragma unroll
        for (int j = 0; j < STRIP + 10; j++) { a[j] = sX[row * ISTRIDE + w * STRIP + j]; b[j] = sY[row * ISTRIDE + w * STRIP + j]; }
        conv_strip(a, o);
#pragma unroll
        for (int j = 0; j < STRIP; j++) hb.v[0][row][w * STRIP + j] = o[j];
        conv_strip(b, o);
#pragma unroll
        for (int j = 0; j < STRIP; j++) hb.v[2][row][w * STRIP + j] = o[j];
#pragma unroll
        for (int j = 0; j < STRIP + 10; j++) p[j] = a[j] * a[j];
        conv_strip(p, o);
#pragma unroll
        for (int j = 0; j < STRIP; j++) hb.v[1][row][w * STRIP + j] = o[j];
#pragma unroll
        for (int j = 0; j < STRIP + 10; j++) p[j] = b[j] * b[j];
        conv_strip(p, o);
#pragma unroll
        for (int j = 0; j < STRIP; j++) hb.v[3][row][w * STRIP + j] = o[j];
#pragma unroll
        for (int j = 0; j < STRIP + 10; j++) p[j] = a[j] * b[j];
        conv_strip(p, o);
#pragma unroll
        for (int j = 0; j < STRIP; j++) hb.v[4][row][w * STRIP + j] = o[j];
    }
    __syncthreads();
    // vertical pass: thread = (column, strip of 8 rows)
    const int cx = t & (TX - 1), ys = (t >> 6) * STRIP;
    float mom[5][STRIP];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        float in[STRIP + 10];
#pragma unroll
        for (int j = 0; j < STRIP + 10; j++) in[j] = hb.v[k][ys + j][cx];
        conv_strip(in, mom[k]);
    }
    const int gx = x0 + cx;
    float lsum = 0.0f;
#pragma unroll
    for (int j = 0; j < STRIP; j++) {
        const int gy = y0 + ys + j;
        if (gx < W && gy < H) {
            float mu1 = mom[0][j], mu2 = mom[2][j];
            float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2;
            float sigma1_sq = mom[1][j] - mu1_sq, sigma2_sq = mom[3][j] - mu2_sq, sigma12 = mom[4][j] - mu1 * mu2;
            float A = mu1_sq + mu2_sq + C1, B = sigma1_sq + sigma2_sq + C2;
            float C = 2.0f * mu1 * mu2 + C1, D = 2.0f * sigma12 + C2;
            float iAB = 1.0f / (A * B);
            float val = (C * D) * iAB;
            float out = val;
            if (L1) {
                float l1 = fabsf(sX[(ys + j + HALO) * ISTRIDE + cx + HALO] - sY[(ys + j + HALO) * ISTRIDE + cx + HALO]);
                out = ssim_weight * (1.0f - val) + (1.0f - ssim_weight) * l1;
            }
            const size_t gi = plane + (size_t)gy * W + gx;
            if (map != nullptr) map[gi] = out;
            lsum += out;
            if (TRAIN) {
                // fused_ssim/ssim.cu:258-268 with the common factor 1/(A B) pulled out
                dm_dmu1[gi] = 2.0f * iAB * (mu2 * (D - C) + mu1 * C * D * (1.0f / B - 1.0f / A));
                dm_dsigma1_sq[gi] = -(C * D) * iAB / B;
                dm_dsigma12[gi] = 2.0f * C * iAB;
            }
        }
    }
    if (block_sums != nullptr) {
        __shared__ float s_part[THREADS / 32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
        if (lane == 0) s_part[w] = lsum;
        __syncthreads();
        if (t == 0) {
            float s = 0.0f;
#pragma unroll
            for (int k = 0; k < THREADS / 32; k++) s += s_part[k];
            block_sums[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = s;
        }
    }
}

// dL_dmap == nullptr: uniform upstream gradient `chain` (loss = mean of the map)
template <bool L1>
__global__ void __launch_bounds__(THREADS) ssim_backward_kernel(const float* __restrict__ img1, const float* __restrict__ img2,
                                                                const float* __restrict__ dL_dmap, float chain,
                                                                const float* __restrict__ dm_dmu1, const float* __restrict__ dm_dsigma1_sq,
                                                                const float* __restrict__ dm_dsigma12, int H, int W, float ssim_weight,
                                                                float* __restrict__ dL_dimg1)
{
    extern __shared__ float smem[];
    float* sM[3] = { smem, smem + IY * ISTRIDE, smem + 2 * IY * ISTRIDE };
    HBuf<3>& hb = *reinterpret_cast<HBuf<3>*>(smem + 3 * IY * ISTRIDE);
    const size_t plane = (size_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const float scale = L1 ? -ssim_weight : 1.0f;                  // d loss / d ssim  (fused_ssim/ssim.cu:775-779)
    for (int i = t; i < IY * IX; i += THREADS) {
        int ly = i / IX, lx = i - ly * IX;
        int gy = y0 - HALO + ly, gx = x0 - HALO + lx;
        float m0 = 0.0f, m1 = 0.0f, m2 = 0.0f;
        if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
            const size_t gi = plane + (size_t)gy * W + gx;
            const float c = scale * (dL_dmap != nullptr ? dL_dmap[gi] : chain);
            m0 = dm_dmu1[gi] * c; m1 = dm_dsigma1_sq[gi] * c; m2 = dm_dsigma12[gi] * c;
        }
        sM[0][ly * ISTRIDE + lx] = m0; sM[1][ly * ISTRIDE + lx] = m1; sM[2][ly * ISTRIDE + lx] = m2;
    }
    __syncthreads();
    for (int row = lane; row < IY; row += 32) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float a[STRIP + 10], o[STRIP];
#pragma unroll
            for (int j = 0; j < STRIP + 10; j++) a[j] = sM[k][row * ISTRIDE + w * STRIP + j];
            conv_strip(a, o);
#pragma unroll
            for (int j = 0; j < STRIP; j++) hb.v[k][row][w * STRIP + j] = o[j];
        }
    }
    __syncthreads();
    const int cx = t & (TX - 1), ys = (t >> 6) * STRIP;
    float s[3][STRIP];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float in[STRIP + 10];
#pragma unroll
        for (int j = 0; j < STRIP + 10; j++) in[j] = hb.v[k][ys + j][cx];
        conv_strip(in, s[k]);
    }
    const int gx = x0 + cx;
#pragma unroll
    for (int j = 0; j < STRIP; j++) {
        const int gy = y0 + ys + j;
        if (gx < W && gy < H) {
            const size_t gi = plane + (size_t)gy * W + gx;
            const float p1 = img1[gi], p2 = img2[gi];
            float g = s[0][j] + (2.0f * p1) * s[1][j] + p2 * s[2][j];             // fused_ssim/ssim.cu:417-421
            if (L1) {
                const float c = dL_dmap != nullptr ? dL_dmap[gi] : chain;
                const float sg = (p1 == p2) ? 0.0f : copysignf(1.0f, p1 - p2);      // fused_ssim/ssim.cu:840-841
                g += (1.0f - ssim_weight) * sg * c;
            }
            dL_dimg1[gi] = g;
        }
    }
}

constexpr size_t FWD_SMEM = (2 * IY * ISTRIDE) * sizeof(float) + sizeof(HBuf<5>);
constexpr size_t BWD_SMEM = (3 * IY * ISTRIDE) * sizeof(float) + sizeof(HBuf<3>);

template <typename K>
int opt_in_smem(K kernel, size_t bytes)
{
    LGS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return LGS_OK;
}

}  // namespace

/* number of per-CTA partial sums lgs_ssim_forward writes for this shape */
extern "C" int lgs_ssim_num_block_sums(int B, int CH, int H, int W, int* count)
{
    LGS_REQUIRE(B >= 1 && CH >= 1 && H >= 1 && W >= 1 && count != nullptr, "ssim_num_block_sums: bad arguments");
    *count = B * CH * ((H + TY - 1) / TY) * ((W + TX - 1) / TX);
    return LGS_OK;
}

extern "C" int lgs_ssim_forward(const float* img1, const float* img2, int B, int CH, int H, int W, float C1, float C2, int l1_mode,
                                float ssim_weight, float* map, float* dm_dmu1, float* dm_dsigma1_sq, float* dm_dsigma12,
                                float* block_sums, void* stream)
{
    LGS_REQUIRE(B >= 1 && CH >= 1 && H >= 1 && W >= 1, "ssim_forward: bad shape [%d,%d,%d,%d]", B, CH, H, W);
    LGS_REQUIRE(img1 != nullptr && img2 != nullptr, "ssim_forward: null image");
    LGS_REQUIRE((size_t)B * CH <= 65535, "ssim_forward: B*CH = %d exceeds the grid z limit", B * CH);
    const bool train = dm_dmu1 != nullptr;
    LGS_REQUIRE(!train || (dm_dsigma1_sq != nullptr && dm_dsigma12 != nullptr), "ssim_forward: the three partial maps come together");
    LGS_REQUIRE(map != nullptr || block_sums != nullptr || train, "ssim_forward: nothing to compute");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((W + TX - 1) / TX, (H + TY - 1) / TY, B * CH);
#define FWD(L, T)                                                                                                              \
    do {                                                                                                                       \
        static int once = opt_in_smem(ssim_forward_kernel<L, T>, FWD_SMEM);                                                    \
        if (once != LGS_OK) return once;                                                                                       \
        ssim_forward_kernel<L, T><<<grid, THREADS, FWD_SMEM, st>>>(img1, img2, H, W, C1, C2, ssim_weight, map, dm_dmu1, dm_dsigma1_sq, \
                                                                  dm_dsigma12, block_sums);                                    \
    } while (0)
    if (l1_mode) { if (train) FWD(true, true); else FWD(true, false); }
    else { if (train) FWD(false, true); else FWD(false, false); }
#undef FWD
    LGS_CHECK_LAUNCH("ssim_forward_kernel");
    return LGS_OK;
}

extern "C" int lgs_ssim_backward(const float* img1, const float* img2, const float* dL_dmap, float uniform_chain, const float* dm_dmu1,
                                 const float* dm_dsigma1_sq, const float* dm_dsigma12, int B, int CH, int H, int W, int l1_mode,
                                 float ssim_weight, float* dL_dimg1, void* stream)
{
    LGS_REQUIRE(B >= 1 && CH >= 1 && H >= 1 && W >= 1, "ssim_backward: bad shape [%d,%d,%d,%d]", B, CH, H, W);
    LGS_REQUIRE(img1 && img2 && dm_dmu1 && dm_dsigma1_sq && dm_dsigma12 && dL_dimg1, "ssim_backward: null tensor");
    LGS_REQUIRE((size_t)B * CH <= 65535, "ssim_backward: B*CH = %d exceeds the grid z limit", B * CH);
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((W + TX - 1) / TX, (H + TY - 1) / TY, B * CH);
#define BWD(L)                                                                                                                 \
    do {                                                                                                                       \
        static int once = opt_in_smem(ssim_backward_kernel<L>, BWD_SMEM);                                                      \
        if (once != LGS_OK) return once;                                                                                       \
        ssim_backward_kernel<L><<<grid, THREADS, BWD_SMEM, st>>>(img1, img2, dL_dmap, uniform_chain, dm_dmu1, dm_dsigma1_sq,   \
                                                                dm_dsigma12, H, W, ssim_weight, dL_dimg1);                     \
    } while (0)
    if (l1_mode) BWD(true); else BWD(false);
#undef BWD
    LGS_CHECK_LAUNCH("ssim_backward_kernel");
    return LGS_OK;
}
