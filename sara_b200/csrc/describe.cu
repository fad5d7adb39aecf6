// Dominant orientations and 128-D SIFT descriptors, one warp per keypoint.
//
// Restates the behaviour of
//   gradient_polar_coordinates      FeatureDescriptors/Orientation.cpp:24-56,
//                                   ImageProcessing/Differential.hpp:46-61
//   compute_orientation_histogram   FeatureDescriptors/Orientation.hpp:91-139
//   lowe_smooth_histogram           Orientation.hpp:147-165
//   find_peaks / refine_peak        Orientation.hpp:176-214
//   ComputeDominantOrientations     Orientation.cpp:90-166
//   ComputeSIFTDescriptor<4, 8>     FeatureDescriptors/SIFT.hpp:62-145, 204-258
//   rescale to image coordinates    FeatureDetectors/SIFT.cpp:92-98
//
// The reference materialises a polar-gradient pyramid (magnitude, atan2) for
// all 6 layers of every octave; here the gradient is evaluated on the fly from
// the Gaussian layer G(s, o) for just the samples a keypoint touches.
//
// Determinism: every lane accumulates into a private shared-memory histogram
// (no floating-point atomics), then the 32 partials of each bin are summed in a
// fixed order, so the same frame always yields the same bits.  Relative to the
// CPU path the SUMMATION ORDER differs (raster order there), and atan2f / expf
// are CUDA's, so orientations and descriptors are compared with a tolerance.
#include "common.cuh"
#include "scan.cuh"

namespace sb {

  namespace {

    constexpr float kPi = 3.14159274101257324f;     // float(M_PI)
    constexpr float kTwoPi = 6.28318548202514648f;  // float(2. * M_PI)

    constexpr int ORI_WARPS = 8;

    __global__ void __launch_bounds__(ORI_WARPS * 32)
        orientation_kernel(const __grid_constant__ PyramidDesc P, const Keypoint* __restrict__ ext,
                           Counters* __restrict__ counters, int cap_ext,
                           int* __restrict__ ori_count, float* __restrict__ oris)
    {
      __shared__ float s_priv[ORI_WARPS][36][32];
      __shared__ float s_hist[ORI_WARPS][2][36];
      const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
      const int n = min(counters->n_ext, cap_ext);
      float(*priv)[32] = s_priv[wid];

      // extrema are handed out through a device-side queue: their cost grows with the scale
      while (true)
      {
        int i = 0;
        if (lane == 0)
          i = atomicAdd(&counters->ori_next, 1);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= n)
          break;
        const Keypoint kp = ext[i];
        const OctaveDesc& oc = P.oct[kp.o];
        const float* G = oc.G + static_cast<size_t>(kp.s) * oc.layer_stride;
        const int w = oc.w, h = oc.h, pitch = oc.pitch;

        // Orientation.cpp:150-157: the DISCRETE scale of the layer, not the refined one.
        const float scale = P.scale_rel[kp.s];
        const float sigma = __fmul_rn(scale, 1.5f);
        const int rx = static_cast<int>(roundf(kp.x));
        const int ry = static_cast<int>(roundf(kp.y));
        const int radius = static_cast<int>(roundf(__fmul_rn(sigma, 3.f)));
        const float denom = __fmul_rn(__fmul_rn(2.f, sigma), sigma);

#pragma unroll
        for (int b = 0; b < 36; ++b)
          priv[b][lane] = 0.f;

        const int side = 2 * radius + 1;
        const int count = side * side;
        const unsigned magic = (1u << 22) / static_cast<unsigned>(side) + 1u;
        // windows that (with the gradient stencil) lie inside the layer skip the border variants of the loads
        const bool interior = rx - radius >= 1 && rx + radius <= w - 2 && ry - radius >= 1 && ry + radius <= h - 2;
        // three samples per lane and round, their gradient loads requested together
        for (int t0 = lane; t0 < count; t0 += 96)
        {
          float xn[3], xp[3], yn[3], yp[3];
          int uu[3], vv[3];
          bool on[3];
#pragma unroll
          for (int k = 0; k < 3; ++k)
          {
            const int t = t0 + 32 * k;
            // t / side by a magic multiply (exact while side^3 < 2^22: radius < 80)
            const int vr = side < 160 ? static_cast<int>((static_cast<unsigned>(t) * magic) >> 22) : t / side;
            vv[k] = vr - radius;
            uu[k] = t - vr * side - radius;
            const int X = rx + uu[k], Y = ry + vv[k];
            on[k] = t < count && X >= 0 && X < w && Y >= 0 && Y < h;
            xn[k] = xp[k] = yn[k] = yp[k] = 0.f;
            if (on[k] && interior)
            {
              const float* p = G + static_cast<size_t>(Y) * pitch + X;
              xn[k] = __ldg(p + 1);
              xp[k] = __ldg(p - 1);
              yn[k] = __ldg(p + pitch);
              yp[k] = __ldg(p - pitch);
            }
            else if (on[k])
            {
              // Gradient functor: central differences, one-sided at the borders (Differential.hpp:46-61)
              const float* row = G + static_cast<size_t>(Y) * pitch;
              xn[k] = __ldg(row + (X == w - 1 ? X : X + 1));
              xp[k] = __ldg(row + (X == 0 ? X : X - 1));
              yn[k] = __ldg(G + static_cast<size_t>(Y == h - 1 ? Y : Y + 1) * pitch + X);
              yp[k] = __ldg(G + static_cast<size_t>(Y == 0 ? Y : Y - 1) * pitch + X);
            }
          }
#pragma unroll
          for (int k = 0; k < 3; ++k)
          {
            if (!on[k])
              continue;
            // "/ 2" as "* 0.5": the same bits (an exact scaling by a power of two), a tenth of the instructions
            const float gx = __fmul_rn(__fsub_rn(xn[k], xp[k]), 0.5f);
            const float gy = __fmul_rn(__fsub_rn(yn[k], yp[k]), 0.5f);
            const float mag = __fmul_rn(2.f, __fsqrt_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy))));
            float ori = atan2f(gy, gx);
            ori = ori < 0.f ? __fadd_rn(ori, kTwoPi) : ori;
            int bin = static_cast<int>(floorf(__fmul_rn(__fdiv_rn(ori, kTwoPi), 36.f)));
            bin %= 36;
            const float weight = expf(__fdiv_rn(static_cast<float>(-(uu[k] * uu[k] + vv[k] * vv[k])), denom));
            priv[bin][lane] = __fadd_rn(priv[bin][lane], __fmul_rn(weight, mag));
          }
        }
        __syncwarp();

        // Fixed-order reduction of the 32 partials of each bin.
        float* h0 = s_hist[wid][0];
        float* h1 = s_hist[wid][1];
        for (int b = lane; b < 36; b += 32)
        {
          float sum = 0.f;
#pragma unroll 8
          for (int j = 0; j < 32; ++j)
            sum = __fadd_rn(sum, priv[b][(j + lane) & 31]);
          h0[b] = sum;
        }
        __syncwarp();

        // lowe_smooth_histogram, 6 iterations: new[i] = ((old[i-1] + old[i]) + old[i+1]) / 3.
        for (int iter = 0; iter < 6; ++iter)
        {
          for (int b = lane; b < 36; b += 32)
          {
            const float prev = h0[b == 0 ? 35 : b - 1];
            const float next = h0[b == 35 ? 0 : b + 1];
            h1[b] = __fdiv_rn(__fadd_rn(__fadd_rn(prev, h0[b]), next), 3.f);
          }
          __syncwarp();
          float* t = h0;
          h0 = h1;
          h1 = t;
        }

        // find_peaks + refine_peak.
        float mx = fmaxf(h0[lane], lane < 4 ? h0[32 + lane] : h0[lane]);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1)
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        const float thres = __fmul_rn(0.8f, mx);
        int n_peaks = 0;
        for (int base = 0; base < 36; base += 32)
        {
          const int b = base + lane;
          bool is_peak = false;
          float y0 = 0.f, y1 = 0.f, y2 = 0.f;
          if (b < 36)
          {
            y0 = h0[b == 0 ? 35 : b - 1];
            y1 = h0[b];
            y2 = h0[b == 35 ? 0 : b + 1];
            is_peak = y1 >= thres && y1 > y0 && y1 > y2;
          }
          const unsigned m = __ballot_sync(0xffffffffu, is_peak);
          if (is_peak)
          {
            const int slot = n_peaks + __popc(m & ((1u << lane) - 1u));
            const float fprime = __fdiv_rn(__fsub_rn(y2, y0), 2.f);
            const float fsecond = __fadd_rn(__fsub_rn(y0, __fmul_rn(2.f, y1)), y2);
            const float hh = __fdiv_rn(-fprime, fsecond);
            float p = __fadd_rn(__fadd_rn(static_cast<float>(b), 0.5f), hh);
            p = __fmul_rn(p, __fdiv_rn(kTwoPi, 36.f));
            if (p > kPi)
              p = __fsub_rn(p, __fmul_rn(2.f, kPi));
            if (slot < kMaxOri)
              oris[static_cast<size_t>(i) * kMaxOri + slot] = p;
          }
          n_peaks += __popc(m);
        }
        if (lane == 0)
          ori_count[i] = min(n_peaks, kMaxOri);
        __syncwarp();
      }
    }

    // Ordered expansion: keypoint i is emitted ori_count[i] times (Orientation.cpp:158-163).
    __global__ void __launch_bounds__(256)
        expand_kernel(const Keypoint* __restrict__ ext, const int* __restrict__ ori_count,
                      const float* __restrict__ oris, const int* __restrict__ off,
                      const int* __restrict__ chunk_off, const Counters* __restrict__ counters,
                      int cap_ext, Keypoint* __restrict__ kp_oct, int cap_kp)
    {
      const int n = min(counters->n_ext, cap_ext);
      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
      {
        const int c = ori_count[i];
        if (c == 0)
          continue;
        Keypoint kp = ext[i];
        const int pos = off[i] + chunk_off[i >> 10];
        for (int j = 0; j < c; ++j)
          if (pos + j < cap_kp)
          {
            kp.orientation = oris[static_cast<size_t>(i) * kMaxOri + j];
            kp_oct[pos + j] = kp;
          }
      }
    }

    constexpr int DESC_WARPS = 4;
    constexpr int DESC_NC = 16;  // copies of the 128-bin histogram per warp; 32 / DESC_NC update phases
    // 128 real bins + 8 scratch bins, one per contribution of a sample: a contribution whose
    // "+1" spatial neighbour does not exist (xi == 3 or yi == 3) goes to ITS scratch bin, so the
    // eight addresses of a sample are always distinct and its eight loads can be issued together
    // (one shared-memory latency per update instead of eight dependent ones).
    constexpr int DESC_BINS = 128 + 8;

    // atan2 for the descriptor's soft orientation binning: odd minimax polynomial on [0, 1]
    // (|error| < 1e-5 rad, i.e. 1.3e-5 of a bin; the binning is continuous in the angle, so this
    // stays far inside the descriptor tolerance).  The dominant-orientation kernel keeps atan2f:
    // its binning is a hard floor().
    __device__ __forceinline__ float fast_atan2(float y, float x)
    {
      const float ax = fabsf(x), ay = fabsf(y);
      const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
      // mn / mx by the approximate reciprocal (gradient differences are far from the denormal range, where
      // the flush-to-zero form would give inf and the guard below 0)
      float rcp;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(mx));
      const float a = mx > 1e-30f ? __fmul_rn(mn, rcp) : 0.f;
      const float s = __fmul_rn(a, a);
      float r = __fmaf_rn(s, -0.0117212f, 0.05265332f);
      r = __fmaf_rn(r, s, -0.11643287f);
      r = __fmaf_rn(r, s, 0.19354346f);
      r = __fmaf_rn(r, s, -0.33262347f);
      r = __fmaf_rn(r, s, 0.99997726f);
      r = __fmul_rn(r, a);
      r = ay > ax ? __fsub_rn(1.57079637f, r) : r;
      r = x < 0.f ? __fsub_rn(kPi, r) : r;
      return y < 0.f ? -r : r;
    }

    __device__ __forceinline__ float approx_sqrt(float v)
    {
      float r;
      asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
      return r;
    }

    // One warp per keypoint, keypoints handed out through a device-side queue (their cost
    // varies with the square of the scale).  The window of ComputeSIFTDescriptor (SIFT.hpp:62-145)
    // is the bounding square of the ROTATED 4x4 grid, so about half of its pixels fall outside
    // the grid: a cheap geometric test runs on all pixels, the survivors are compacted through a
    // small per-warp queue (ballot + popc), and the expensive part (gradient, atan2, exp,
    // trilinear update of 8 bins) runs on full warps.
    __global__ void __launch_bounds__(DESC_WARPS * 32, 5)
        descriptor_kernel(const __grid_constant__ PyramidDesc P, const Keypoint* __restrict__ kp_oct,
                          Counters* __restrict__ counters, int cap_kp, Keypoint* __restrict__ kp_out,
                          float* __restrict__ desc)
    {
      extern __shared__ float s_dyn[];  // DESC_WARPS x ((128 + 8) bins x 16 copies + 128 queue entries)
      const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
      float* priv = s_dyn + wid * (DESC_BINS * DESC_NC + 128);
      int* queue = reinterpret_cast<int*>(priv + DESC_BINS * DESC_NC);
      const int lcopy = lane & (DESC_NC - 1);
      const int lphase = lane / DESC_NC;
      const int n = min(counters->n_kp, cap_kp);

      while (true)
      {
        int i = 0;
        if (lane == 0)
          i = atomicAdd(&counters->desc_next, 1);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= n)
          break;
        const Keypoint kp = kp_oct[i];
        const OctaveDesc& oc = P.oct[kp.o];
        const float* G = oc.G + static_cast<size_t>(kp.s) * oc.layer_stride;
        const int w = oc.w, h = oc.h, pitch = oc.pitch;

        // OERegion::scale() -> radius(), Features/Feature.cpp:28-39, for shape = a * I.
        const float rr = __fdiv_rn(1.f, __fsqrt_rn(kp.shape[0]));
        const float s = __fsqrt_rn(__fadd_rn(__fmul_rn(rr, rr), 0.f));
        const float theta = kp.orientation;
        const float l = __fmul_rn(3.f, s);
        const float r = __fdiv_rn(__fmul_rn(__fmul_rn(__fsqrt_rn(2.f), l), 5.f), 2.f);
        // cosf / sinf evaluated in double and rounded: agrees with a correctly rounded libm.
        const float ct = static_cast<float>(cos(static_cast<double>(theta)));
        const float st = static_cast<float>(sin(static_cast<double>(theta)));
        const float T00 = __fdiv_rn(ct, l), T01 = __fdiv_rn(st, l);
        const float T10 = __fdiv_rn(-st, l), T11 = __fdiv_rn(ct, l);
        const int rounded_r = static_cast<int>(roundf(r));
        const int rx = static_cast<int>(roundf(kp.x));
        const int ry = static_cast<int>(roundf(kp.y));

        // Histogram: DESC_NC copies of the 128 bins; the lanes that share a copy update it in
        // separate phases (fixed order => deterministic sums, no bank conflicts).
#pragma unroll 4
        for (int b = 0; b < DESC_BINS * DESC_NC / 32; ++b)
          priv[b * 32 + lane] = 0.f;

        // A kept sample in two steps, so that the loads of several samples are in flight
        // together: fetch() issues the four gradient loads, finish() does the trilinear
        // accumulate() (SIFT.hpp:204-238) into the lane's private bins.
        struct Sample
        {
          float fu, fv, xn, xp, yn, yp;
        };
        // Keypoints whose whole window (plus the gradient stencil) lies inside the layer skip
        // the one-sided border differences of the Gradient functor (Differential.hpp:46-61).
        const bool interior = rx - rounded_r >= 1 && rx + rounded_r <= w - 2 && ry - rounded_r >= 1 &&
                              ry + rounded_r <= h - 2;
        const float* Gc = G + static_cast<size_t>(ry) * pitch + rx;
        auto fetch = [&](int uv, bool on) {
          Sample sm;
          const int u = static_cast<short>(uv & 0xffff), v = uv >> 16;
          sm.fu = static_cast<float>(u);
          sm.fv = static_cast<float>(v);
          sm.xn = sm.xp = sm.yn = sm.yp = 0.f;
          if (on)
          {
            if (interior)
            {
              const float* p = Gc + (v * pitch + u);
              sm.xn = __ldg(p + 1);
              sm.xp = __ldg(p - 1);
              sm.yn = __ldg(p + pitch);
              sm.yp = __ldg(p - pitch);
            }
            else
            {
              const int X = rx + u, Y = ry + v;
              const float* row = G + static_cast<size_t>(Y) * pitch;
              const int xn = X == w - 1 ? X : X + 1, xp = X == 0 ? X : X - 1;
              const int yn = Y == h - 1 ? Y : Y + 1, yp = Y == 0 ? Y : Y - 1;
              sm.xn = __ldg(row + xn);
              sm.xp = __ldg(row + xp);
              sm.yn = __ldg(G + static_cast<size_t>(yn) * pitch + X);
              sm.yp = __ldg(G + static_cast<size_t>(yp) * pitch + X);
            }
          }
          return sm;
        };
        // What one kept sample adds to the histogram: 8 (address, value) pairs with DISTINCT
        // addresses (a missing +1 neighbour -- xi == 3 or yi == 3 -- is redirected to the
        // contribution's own scratch bin).
        struct Weights
        {
          int a[8];  // float offsets into the lane's histogram copy
          float v[8];
        };
        // ... in two steps: geom() is everything that depends on the sample's POSITION only and runs for all
        // three samples of a lane while their gradient loads are in flight; finish() needs the loaded values.
        struct Geom
        {
          float weight, xfrac, yfrac, a00, a01, a10, a11;
          int cell;
          bool x1, y1;
        };
        auto geom = [&](const Sample& sm) {
          Geom g;
          float px = __fadd_rn(__fmul_rn(T00, sm.fu), __fmul_rn(T01, sm.fv));
          float py = __fadd_rn(__fmul_rn(T10, sm.fu), __fmul_rn(T11, sm.fv));
          // exp(-(px^2 + py^2) / 8) = 2^(-(px^2 + py^2) * log2(e) / 8); the argument stays above -3, no flush needed
          asm("ex2.approx.ftz.f32 %0, %1;"
              : "=f"(g.weight)
              : "f"(__fmul_rn(__fadd_rn(__fmul_rn(px, px), __fmul_rn(py, py)), -0.180336880f)));
          px = __fadd_rn(px, 1.5f);
          py = __fadd_rn(py, 1.5f);
          // std::modf truncates toward zero (quirk N6): for pos in (-1, 0) the integer part is 0,
          // the weight of cell 0 is 1 - frac > 1 and the one of cell 1 is frac < 0.
          const int xi = static_cast<int>(px), yi = static_cast<int>(py);
          g.xfrac = __fsub_rn(px, static_cast<float>(xi));
          g.yfrac = __fsub_rn(py, static_cast<float>(yi));
          // (xi, yi) in [0, 3]; the +1 neighbours exist for xi, yi < 3
          const float wy0 = __fsub_rn(1.f, g.yfrac), wx0 = __fsub_rn(1.f, g.xfrac);
          g.x1 = xi < 3;
          g.y1 = yi < 3;
          g.cell = ((4 * yi + xi) * 8) * DESC_NC + lcopy;
          g.a00 = __fmul_rn(wy0, wx0);
          g.a01 = __fmul_rn(wy0, g.xfrac);
          g.a10 = __fmul_rn(g.yfrac, wx0);
          g.a11 = __fmul_rn(g.yfrac, g.xfrac);
          return g;
        };
        auto finish = [&](const Sample& sm, const Geom& g) {
          Weights w;
          const float dx = __fsub_rn(sm.xn, sm.xp), dy = __fsub_rn(sm.yn, sm.yp);
          const float mag = approx_sqrt(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));  // = 2 |(dx, dy) / 2|
          float ori = __fsub_rn(fast_atan2(dy, dx), theta);
          ori = ori < 0.f ? __fadd_rn(ori, kTwoPi) : ori;
          ori = __fmul_rn(ori, 1.27323954f);  // 8 / (2 pi)
          const int oi = static_cast<int>(ori);
          const float ofrac = __fsub_rn(ori, static_cast<float>(oi));
          const float wm = __fmul_rn(g.weight, mag);
          const float wo1 = __fmul_rn(ofrac, wm), wo0 = __fsub_rn(wm, wo1);
          const bool x1 = g.x1, y1 = g.y1;
          const int cell = g.cell;
          const int o0 = (oi & 7) * DESC_NC, o1 = ((oi + 1) & 7) * DESC_NC;
          const int cx = cell + 8 * DESC_NC, cy = cell + 32 * DESC_NC;
          const int scratch = 128 * DESC_NC + lcopy;
          w.a[0] = cell + o0;
          w.a[1] = cell + o1;
          w.a[2] = x1 ? cx + o0 : scratch + 2 * DESC_NC;
          w.a[3] = x1 ? cx + o1 : scratch + 3 * DESC_NC;
          w.a[4] = y1 ? cy + o0 : scratch + 4 * DESC_NC;
          w.a[5] = y1 ? cy + o1 : scratch + 5 * DESC_NC;
          w.a[6] = (x1 && y1) ? cy + 8 * DESC_NC + o0 : scratch + 6 * DESC_NC;
          w.a[7] = (x1 && y1) ? cy + 8 * DESC_NC + o1 : scratch + 7 * DESC_NC;
          w.v[0] = __fmul_rn(g.a00, wo0);
          w.v[1] = __fmul_rn(g.a00, wo1);
          w.v[2] = __fmul_rn(g.a01, wo0);
          w.v[3] = __fmul_rn(g.a01, wo1);
          w.v[4] = __fmul_rn(g.a10, wo0);
          w.v[5] = __fmul_rn(g.a10, wo1);
          w.v[6] = __fmul_rn(g.a11, wo0);
          w.v[7] = __fmul_rn(g.a11, wo1);
          return w;
        };
        // the 8 bin updates of one sample, for the lanes whose turn it is: all loads, then all stores
        auto update = [&](const Weights& w, bool mine) {
          if (!mine)
            return;
          float cur[8];
#pragma unroll
          for (int k = 0; k < 8; ++k)
            cur[k] = priv[w.a[k]];
#pragma unroll
          for (int k = 0; k < 8; ++k)
            priv[w.a[k]] = __fadd_rn(cur[k], w.v[k]);
        };
        // Processes queue entries [0, m), m <= 96, three per lane with their loads overlapped.
        auto drain = [&](int m) {
          const bool on0 = lane < m, on1 = lane + 32 < m, on2 = lane + 64 < m;
          const Sample a = fetch(queue[lane], on0);
          const Sample b = fetch(queue[lane + 32], on1);
          const Sample c = fetch(queue[lane + 64], on2);
          const Geom ga = geom(a), gb = geom(b), gc = geom(c);
          const Weights wa = finish(a, ga), wb = finish(b, gb), wc = finish(c, gc);
#pragma unroll
          for (int ph = 0; ph < 32 / DESC_NC; ++ph)
          {
            update(wa, on0 && lphase == ph);
            update(wb, on1 && lphase == ph);
            update(wc, on2 && lphase == ph);
            __syncwarp();
          }
        };

        // The kept pixels of the bounding square [-rounded_r, rounded_r]^2, in raster order.  The test of
        // SIFT.hpp:84-110 (-1 < pos < 4 on both axes, pos = T (u, v) + 1.5) is, in float arithmetic too, a
        // conjunction of four conditions that are MONOTONE in u (every rounded operation is monotone), so on a
        // row the kept pixels form exactly one interval.  A lane works out the interval of one row: first from
        // the inverse inequalities, widened by a pixel on both sides, then it moves both ends inwards until the
        // exact float test accepts them.  The warp then copies the intervals into its queue 32 pixels at a
        // time -- no per-pixel test, no ballot -- and about half of the square is never visited.
        int q_n = 0;  // entries waiting in the queue (< 96 between rounds)
        const int u_min = max(-rounded_r, -rx), u_max = min(rounded_r, w - 1 - rx);
        for (int v0 = -rounded_r; v0 <= rounded_r; v0 += 32)
        {
          int my_lo = 1, my_hi = 0;  // empty
          {
            const int v = v0 + lane, Y = ry + v;
            if (v <= rounded_r && Y >= 0 && Y < h)
            {
              const float fv = static_cast<float>(v);
              const float tx = __fmul_rn(T01, fv), ty = __fmul_rn(T11, fv);
              const float bx = tx + 1.5f, by = ty + 1.5f;
              float lo = -1e9f, hi = 1e9f;
              if (fabsf(T00) > 1e-12f)
              {
                const float e0 = (-1.f - bx) / T00, e1 = (4.f - bx) / T00;
                lo = fmaxf(lo, fminf(e0, e1));
                hi = fminf(hi, fmaxf(e0, e1));
              }
              if (fabsf(T10) > 1e-12f)
              {
                const float e0 = (-1.f - by) / T10, e1 = (4.f - by) / T10;
                lo = fmaxf(lo, fminf(e0, e1));
                hi = fminf(hi, fmaxf(e0, e1));
              }
              lo = fmaxf(lo, -70000.f);
              hi = fminf(hi, 70000.f);
              if (lo <= hi)
              {
                my_lo = max(u_min, static_cast<int>(floorf(lo)) - 1);
                my_hi = min(u_max, static_cast<int>(ceilf(hi)) + 1);
                // the reference's test, operation for operation
                auto kept = [&](int u) {
                  const float fu = static_cast<float>(u);
                  const float px = __fadd_rn(__fadd_rn(__fmul_rn(T00, fu), tx), 1.5f);
                  const float py = __fadd_rn(__fadd_rn(__fmul_rn(T10, fu), ty), 1.5f);
                  return fminf(px, py) > -1.f && fmaxf(px, py) < 4.f;
                };
                while (my_lo <= my_hi && !kept(my_lo))
                  ++my_lo;
                while (my_lo <= my_hi && !kept(my_hi))
                  --my_hi;
              }
            }
          }
          const int n_rows = min(32, rounded_r - v0 + 1);
          for (int rrow = 0; rrow < n_rows; ++rrow)
          {
            const int ulo = __shfl_sync(0xffffffffu, my_lo, rrow), uhi = __shfl_sync(0xffffffffu, my_hi, rrow);
            const int vhi = (v0 + rrow) << 16;
            for (int ub = ulo; ub <= uhi; ub += 32)
            {
              const int u = ub + lane;
              if (u <= uhi)
                queue[q_n + lane] = (u & 0xffff) | vhi;
              q_n += min(32, uhi - ub + 1);
              __syncwarp();
              if (q_n >= 96)
              {
                drain(96);
                // move the leftover (< 32 entries) to the front
                const int rest = q_n - 96;
                const int moved = lane < rest ? queue[96 + lane] : 0;
                __syncwarp();
                if (lane < rest)
                  queue[lane] = moved;
                q_n = rest;
                __syncwarp();
              }
            }
          }
        }
        drain(q_n);
        __syncwarp();

        // Fixed-order reduction: lane owns bins lane, lane+32, lane+64, lane+96.
        float hv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
          const float* pb = priv + (q * 32 + lane) * DESC_NC;
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < DESC_NC; ++j)
            sum = __fadd_rn(sum, pb[(j + lane) & (DESC_NC - 1)]);
          hv[q] = sum;
        }
        __syncwarp();

        // normalize(): L2, clamp at 0.2, L2 (SIFT.hpp:241-252); then * 512, min 255 (SIFT.hpp:128).
#pragma unroll
        for (int pass = 0; pass < 2; ++pass)
        {
          float sq = __fadd_rn(__fadd_rn(__fmul_rn(hv[0], hv[0]), __fmul_rn(hv[1], hv[1])),
                               __fadd_rn(__fmul_rn(hv[2], hv[2]), __fmul_rn(hv[3], hv[3])));
#pragma unroll
          for (int d = 16; d > 0; d >>= 1)
            sq = __fadd_rn(sq, __shfl_xor_sync(0xffffffffu, sq, d));
          const float nrm = __fsqrt_rn(sq);
#pragma unroll
          for (int q = 0; q < 4; ++q)
          {
            hv[q] = __fdiv_rn(hv[q], nrm);
            if (pass == 0)
              hv[q] = fminf(hv[q], 0.2f);
          }
        }
        float* out = desc + static_cast<size_t>(i) * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          out[q * 32 + lane] = fminf(__fmul_rn(hv[q], 512.f), 255.f);

        if (lane == 0)
        {
          // SIFT.cpp:92-98: back to image coordinates.
          Keypoint f = kp;
          const float z = oc.scaling;
          f.x = __fmul_rn(f.x, z);
          f.y = __fmul_rn(f.y, z);
          const float z2 = __fmul_rn(z, z);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            f.shape[j] = __fdiv_rn(f.shape[j], z2);
          kp_out[i] = f;
        }
      }
    }

  }  // namespace

  int launch_orientation(const PyramidDesc& P, const Keypoint* ext, int cap_ext, int* ori_count,
                         float* oris, int* scratch, Keypoint* kp_oct, int cap_kp, Counters* counters,
                         cudaStream_t st)
  {
    int* chunk_off = scratch;
    int* ori_off = scratch + 1024;
    orientation_kernel<<<148 * 4, ORI_WARPS * 32, 0, st>>>(P, ext, counters, cap_ext, ori_count, oris);
    int launches = 1;
    launches += exclusive_scan(ori_count, ori_off, chunk_off, 0, &counters->n_ext, cap_ext,
                               &counters->n_kp, cap_kp, &counters->overflow, 4, st);
    expand_kernel<<<296, 256, 0, st>>>(ext, ori_count, oris, ori_off, chunk_off, counters, cap_ext, kp_oct,
                                       cap_kp);
    return launches + 1;
  }

  int launch_descriptors(const PyramidDesc& P, const Keypoint* kp_oct, Keypoint* kp_out, float* desc,
                         int cap_kp, Counters* counters, cudaStream_t st)
  {
    const int smem = DESC_WARPS * (DESC_BINS * DESC_NC + 128) * sizeof(float);
    cudaFuncSetAttribute(descriptor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    descriptor_kernel<<<148 * 6, DESC_WARPS * 32, smem, st>>>(P, kp_oct, counters, cap_kp, kp_out, desc);
    return 1;
  }

}  // namespace sb
