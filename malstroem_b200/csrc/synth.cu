// synth.cu — the synthetic fractal DEM of malstroem_b200/synth.py on the device (bit-identical: integer
// value-noise fBm, one float multiply at the end).  Benchmark / test input, not part of the hot path.
#include "common.cuh"

namespace ms {

struct SynthAmpl { long long a[10]; };

__device__ inline long long synth_hash(long long ix, long long iy, uint32_t salt) {
    uint32_t h = ((uint32_t)ix * 0x9E3779B1u) ^ ((uint32_t)iy * 0x85EBCA77u);
    h ^= salt;
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 12;
    h *= 0x297A2D39u;
    h ^= h >> 15;
    return (long long)(h >> 16);
}

__global__ void __launch_bounds__(256) k_synth(float *dem, int rows, int cols, long long row0, long long col0,
                                               int seed, SynthAmpl am) {
    int c = blockIdx.x * 64 + (threadIdx.x & 63);
    int r = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (r >= rows || c >= cols) return;
    long long x = col0 + c, y = row0 + r, total = 0;
#pragma unroll
    for (int o = 0; o < 10; o++) {
        int sh = 10 - o;
        long long ix = x >> sh, iy = y >> sh;
        long long tx = ((x & ((1ll << sh) - 1)) << 16) >> sh;
        long long ty = ((y & ((1ll << sh) - 1)) << 16) >> sh;
        long long sx = (((tx * tx) >> 16) * ((3ll << 16) - 2 * tx)) >> 16;
        long long sy = (((ty * ty) >> 16) * ((3ll << 16) - 2 * ty)) >> 16;
        uint32_t salt = (uint32_t)((unsigned long long)seed * 0xC2B2AE3Dull + (unsigned long long)o * 0x27D4EB2Full);
        long long h00 = synth_hash(ix, iy, salt), h10 = synth_hash(ix + 1, iy, salt);
        long long h01 = synth_hash(ix, iy + 1, salt), h11 = synth_hash(ix + 1, iy + 1, salt);
        long long top = h00 + (((h10 - h00) * sx) >> 16);
        long long bot = h01 + (((h11 - h01) * sx) >> 16);
        long long v = top + (((bot - top) * sy) >> 16);
        total += v * am.a[o];
    }
    long long mm = (total * 200000ll) >> 32;
    dem[(size_t)r * cols + c] = __fmul_rn((float)mm, 0.001f);
}

}  // namespace ms

extern "C" int ms_synth_fractal_dev(float *dem, int64_t rows, int64_t cols, int64_t row0, int64_t col0, int seed,
                                    void *stream) {
    MS_TRY(ms::ensure_init());
    if (!dem || rows < 1 || cols < 1 || rows * cols > (1ll << 32)) { ms::set_error("synth: bad argument"); return MS_ERR_ARG; }
    ms::SynthAmpl am;
    long long a[10], tot = 0;
    a[0] = 1 << 16;
    for (int i = 1; i < 10; i++) a[i] = (a[i - 1] * 40342) >> 16;
    for (int i = 0; i < 10; i++) tot += a[i];
    for (int i = 0; i < 10; i++) am.a[i] = (a[i] << 16) / tot;
    dim3 g2(ms::cdiv(cols, 64), ms::cdiv(rows, 4));
    MS_LAUNCH(ms::k_synth, g2, 256, 0, (cudaStream_t)stream, dem, (int)rows, (int)cols, (long long)row0,
              (long long)col0, seed, am);
    return MS_OK;
}
