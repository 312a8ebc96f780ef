// Integer-pipe microbenchmark for B200 (sm_100a).
// Measures lane-ops/s of the instructions the banded-DP kernel is built from, so that
// (a) the roofline denominator P_int (peak 32-bit integer lane-ops/s) is MEASURED, and
// (b) kernel design choices (16x2 DPX vs scalar, IADD3 vs IMAD, LDS/SHFL/REDUX cost) are evidence-based.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o int_peak int_peak.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <string>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;   // loop trips
constexpr int CH = 8;         // independent chains per thread
constexpr int UNR = 8;        // ops per chain per trip

// Each OP(x, y, z) must be `x = f(x, y, z)` pinned by inline PTX.
#define DEFKERNEL(NAME, OPSTMT)                                                     \
__global__ void __launch_bounds__(512) k_##NAME(uint32_t *out, uint32_t seed) {    \
    uint32_t x[CH];                                                                 \
    uint32_t y = seed * 2654435761u + threadIdx.x, z = seed ^ 0x01010101u;          \
    _Pragma("unroll") for (int c = 0; c < CH; ++c) x[c] = seed + c * 77u + threadIdx.x; \
    for (int it = 0; it < ITERS; ++it) {                                            \
        _Pragma("unroll") for (int u = 0; u < UNR; ++u) {                           \
            _Pragma("unroll") for (int c = 0; c < CH; ++c) { uint32_t &X = x[c]; OPSTMT; } \
        }                                                                           \
    }                                                                               \
    uint32_t acc = 0;                                                               \
    _Pragma("unroll") for (int c = 0; c < CH; ++c) acc ^= x[c];                     \
    if (acc == 0x12345678u) out[threadIdx.x] = acc;                                 \
}

DEFKERNEL(iadd3,      asm volatile("add.u32 %0, %0, %1;" : "+r"(X) : "r"(y)))
DEFKERNEL(iadd3_3in,  asm volatile("{.reg .u32 t; add.u32 t, %0, %1; sub.u32 %0, t, %2;}" : "+r"(X) : "r"(y), "r"(z)))
DEFKERNEL(lop3,       asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(X) : "r"(y), "r"(z)))
DEFKERNEL(shf,        asm volatile("shf.l.wrap.b32 %0, %1, %0, 8;" : "+r"(X) : "r"(y)))
DEFKERNEL(prmt,       asm volatile("prmt.b32 %0, %0, %1, 0x5432;" : "+r"(X) : "r"(y)))
DEFKERNEL(vimnmx,     asm volatile("max.s32 %0, %0, %1;" : "+r"(X) : "r"(y)))
DEFKERNEL(vimnmx3,    X = __vimax3_s32((int)X, (int)y, (int)z))
DEFKERNEL(vimnmx16x2, X = __vmaxs2(X, y))
DEFKERNEL(vimnmx3_16x2, X = __vimax3_s16x2(X, y, z))
DEFKERNEL(viaddmnmx16x2, X = __viaddmax_s16x2_relu(X, y, z))
DEFKERNEL(viadd16x2,  X = __vadd2(X, y))
// UN-FUSABLE 2-input chains: ptxas folds two dependent 2-input adds / max into ONE 3-input IADD3 / VIMNMX3 (SASS of k_iadd3,
// k_vimnmx, k_vimnmx16x2: 512 instructions for 1024 PTX ops), which is what reads as "127 lanes/clk/SM" above.  Alternating a
// 2-input op with an XOR (LOP3, same pipe, cannot be folded into either) shows the true issue rate of the 2-input forms.
#define ALT2(OPA) do { if (u & 1) asm volatile("xor.b32 %0, %0, %1;" : "+r"(X) : "r"(z)); else { OPA; } } while (0)
DEFKERNEL(iadd_nofuse,      ALT2(asm volatile("add.u32 %0, %0, %1;" : "+r"(X) : "r"(y))))
DEFKERNEL(vimnmx_nofuse,    ALT2(asm volatile("max.s32 %0, %0, %1;" : "+r"(X) : "r"(y))))
DEFKERNEL(vimnmx16x2_nofuse, ALT2(X = __vmaxs2(X, y)))
DEFKERNEL(viadd16x2_nofuse, ALT2(X = __vadd2(X, y)))
DEFKERNEL(imad,       asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(X) : "r"(y), "r"(z)))
DEFKERNEL(imad_imm,   asm volatile("mad.lo.u32 %0, %0, 5, %1;" : "+r"(X) : "r"(y)))
DEFKERNEL(imad_hi,    asm volatile("mad.hi.u32 %0, %1, 65536, %0;" : "+r"(X) : "r"(y)))
DEFKERNEL(idp2a,      X = __dp2a_lo(y, 0x00000001u, X))
DEFKERNEL(idp4a,      X = __dp4a(y, 0x00000001u, X))
DEFKERNEL(isetp_sel,  asm volatile("{.reg .pred p; setp.gt.s32 p, %0, %1; selp.u32 %0, %2, %0, p;}" : "+r"(X) : "r"(y), "r"(z)))
DEFKERNEL(shfl,       asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(X) : "r"(z & 31)))
DEFKERNEL(redux,      X = __reduce_max_sync(0xffffffffu, (int)X) + y)
DEFKERNEL(vote,       X = __ballot_sync(0xffffffffu, X > y) + y)
// dual-pipe mix: half the chains use IADD3 (alu pipe), half IMAD.IADD (fma pipe)
__global__ void __launch_bounds__(512) k_mix_alu_fma(uint32_t *out, uint32_t seed) {
    uint32_t x[CH];
    uint32_t y = seed * 2654435761u + threadIdx.x;
    #pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = seed + c * 77u + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int u = 0; u < UNR; ++u) {
            #pragma unroll
            for (int c = 0; c < CH; ++c) {
                if (c & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(y), "r"(seed));
                else       asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y), "r"(seed));
            }
        }
    }
    uint32_t acc = 0;
    #pragma unroll
    for (int c = 0; c < CH; ++c) acc ^= x[c];
    if (acc == 0x12345678u) out[threadIdx.x] = acc;
}
// mix of VIMNMX.S16x2 (alu) + IMAD (fma)
__global__ void __launch_bounds__(512) k_mix_dpx_fma(uint32_t *out, uint32_t seed) {
    uint32_t x[CH];
    uint32_t y = seed * 2654435761u + threadIdx.x;
    #pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = seed + c * 77u + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int u = 0; u < UNR; ++u) {
            #pragma unroll
            for (int c = 0; c < CH; ++c) {
                if (c & 1) x[c] = __viaddmax_s16x2_relu(x[c], y, seed);
                else       asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(y), "r"(seed));
            }
        }
    }
    uint32_t acc = 0;
    #pragma unroll
    for (int c = 0; c < CH; ++c) acc ^= x[c];
    if (acc == 0x12345678u) out[threadIdx.x] = acc;
}
// shared-memory load throughput (LDS.32 and LDS.128), conflict-free
__global__ void __launch_bounds__(512) k_lds32(uint32_t *out, uint32_t seed) {
    __shared__ uint32_t sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = (i * 7 + seed) & 4095;
    __syncthreads();
    uint32_t x[CH];
    #pragma unroll
    for (int c = 0; c < CH; ++c) x[c] = (threadIdx.x + c * 512) & 4095;
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int u = 0; u < UNR; ++u) {
            #pragma unroll
            for (int c = 0; c < CH; ++c) x[c] = sm[x[c]];
        }
    }
    uint32_t acc = 0;
    #pragma unroll
    for (int c = 0; c < CH; ++c) acc ^= x[c];
    if (acc == 0x12345678u) out[threadIdx.x] = acc;
}
__global__ void __launch_bounds__(512) k_lds128(uint32_t *out, uint32_t seed) {
    __shared__ uint4 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_uint4(i, seed, i ^ seed, 0);
    __syncthreads();
    uint4 acc4 = make_uint4(0, 0, 0, 0);
    uint32_t idx = threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
        #pragma unroll
        for (int u = 0; u < UNR * CH; ++u) {
            uint4 v = sm[(idx + u * 512) & 2047];
            acc4.x ^= v.x; acc4.y += v.y; acc4.z ^= v.z; acc4.w += v.w;
        }
        idx += acc4.w & 1;
    }
    uint32_t acc = acc4.x ^ acc4.y ^ acc4.z ^ acc4.w;
    if (acc == 0x12345678u) out[threadIdx.x] = acc;
}

struct Result { std::string name; double glaneops; double ms; };

template <typename K>
static Result run(const char *name, K kernel, int nsm, uint32_t *dout, double ops_per_thread) {
    const int threads = 512, blocks = nsm * 4;   // 2048 threads/SM resident
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) kernel<<<blocks, threads>>>(dout, 1234u + w);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        kernel<<<blocks, threads>>>(dout, 99u + rep);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    double total = ops_per_thread * (double)threads * blocks;
    Result r{ name, total / (best * 1e-3) / 1e9, best };
    return r;
}

int main(int argc, char **argv) {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    int nsm = p.multiProcessorCount;
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
    uint32_t *dout; CK(cudaMalloc(&dout, 4096));
    const double OPS = (double)ITERS * UNR * CH;
    std::vector<Result> rs;
#define RUN(NAME, MULT) rs.push_back(run(#NAME, k_##NAME, nsm, dout, OPS * (MULT)))
    RUN(iadd3, 1); RUN(iadd3_3in, 1); RUN(lop3, 1); RUN(shf, 1); RUN(prmt, 1);
    RUN(vimnmx, 1); RUN(vimnmx3, 1); RUN(vimnmx16x2, 1); RUN(vimnmx3_16x2, 1); RUN(viaddmnmx16x2, 1); RUN(viadd16x2, 1);
    RUN(iadd_nofuse, 1); RUN(vimnmx_nofuse, 1); RUN(vimnmx16x2_nofuse, 1); RUN(viadd16x2_nofuse, 1);
    RUN(imad, 1); RUN(imad_imm, 1); RUN(imad_hi, 1); RUN(idp2a, 1); RUN(idp4a, 1);
    RUN(isetp_sel, 1); RUN(shfl, 1); RUN(redux, 1); RUN(vote, 1);
    RUN(mix_alu_fma, 1); RUN(mix_dpx_fma, 1); RUN(lds32, 1); RUN(lds128, 1);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, \"results\": {", p.name, nsm, clk_khz);
    for (size_t i = 0; i < rs.size(); ++i)
        printf("%s\"%s\": {\"glaneops_per_s\": %.1f, \"ms\": %.4f, \"lanes_per_clk_per_sm_at_max_clk\": %.2f}",
               i ? ", " : "", rs[i].name.c_str(), rs[i].glaneops, rs[i].ms,
               rs[i].glaneops * 1e9 / ((double)nsm * clk_khz * 1e3));
    printf("}}\n");
    return 0;
}
