// microbench.cu -- per-SM throughput of the instructions the HSFM pair loop is made of (B200, sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

template <int OP> __global__ void k(float *out, int iters)
{
    float a0 = threadIdx.x * 1e-3f + 1.0f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    __shared__ float2 sm[320];
    sm[threadIdx.x & 255] = make_float2(a0, a1);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (OP == 0) {   // MUFU.RSQ
                asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a1));
                asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a3));
                asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a4)); asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a5));
                asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a6)); asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a7));
            } else if (OP == 1) {   // MUFU.EX2
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a4)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a5));
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a6)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a7));
            } else if (OP == 2) {   // MUFU.RCP
                asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a1));
                asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a3));
                asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a4)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a5));
                asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a6)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a7));
            } else if (OP == 3) {   // FFMA
                a0 = fmaf(a0, 0.999f, 1e-3f); a1 = fmaf(a1, 0.999f, 1e-3f); a2 = fmaf(a2, 0.999f, 1e-3f); a3 = fmaf(a3, 0.999f, 1e-3f);
                a4 = fmaf(a4, 0.999f, 1e-3f); a5 = fmaf(a5, 0.999f, 1e-3f); a6 = fmaf(a6, 0.999f, 1e-3f); a7 = fmaf(a7, 0.999f, 1e-3f);
            } else if (OP == 4) {   // FFMA with register operands
                a0 = fmaf(a0, a1, a2); a1 = fmaf(a1, a2, a3); a2 = fmaf(a2, a3, a4); a3 = fmaf(a3, a4, a5);
                a4 = fmaf(a4, a5, a6); a5 = fmaf(a5, a6, a7); a6 = fmaf(a6, a7, a0); a7 = fmaf(a7, a0, a1);
            } else if (OP == 5) {   // FFMA2 (4 packed = 8 fma)
                float2 p0 = make_float2(a0, a1), p1 = make_float2(a2, a3), p2 = make_float2(a4, a5), p3 = make_float2(a6, a7);
                const float2 c = make_float2(0.999f, 0.999f), d = make_float2(1e-3f, 1e-3f);
                p0 = __ffma2_rn(p0, c, d); p1 = __ffma2_rn(p1, c, d); p2 = __ffma2_rn(p2, c, d); p3 = __ffma2_rn(p3, c, d);
                p0 = __ffma2_rn(p0, c, d); p1 = __ffma2_rn(p1, c, d); p2 = __ffma2_rn(p2, c, d); p3 = __ffma2_rn(p3, c, d);
                a0 = p0.x; a1 = p0.y; a2 = p1.x; a3 = p1.y; a4 = p2.x; a5 = p2.y; a6 = p3.x; a7 = p3.y;
            } else if (OP == 6) {   // SHFL
                a0 = __shfl_sync(0xffffffffu, a0, (lane + 2) & 31); a1 = __shfl_sync(0xffffffffu, a1, (lane + 2) & 31);
                a2 = __shfl_sync(0xffffffffu, a2, (lane + 2) & 31); a3 = __shfl_sync(0xffffffffu, a3, (lane + 2) & 31);
                a4 = __shfl_sync(0xffffffffu, a4, (lane + 2) & 31); a5 = __shfl_sync(0xffffffffu, a5, (lane + 2) & 31);
                a6 = __shfl_sync(0xffffffffu, a6, (lane + 2) & 31); a7 = __shfl_sync(0xffffffffu, a7, (lane + 2) & 31);
            } else if (OP == 7) {   // LDS.64 conflict-free
                const float2 *p = sm + lane + (i & 1);
                float2 v0, v1, v2, v3, v4, v5, v6, v7;
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v0.x), "=f"(v0.y) : "l"(p));
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+256];" : "=f"(v1.x), "=f"(v1.y) : "l"(p));
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+512];" : "=f"(v2.x), "=f"(v2.y) : "l"(p));
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+768];" : "=f"(v3.x), "=f"(v3.y) : "l"(p));
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+1024];" : "=f"(v4.x), "=f"(v4.y) : "l"(p));
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+1280];" : "=f"(v5.x), "=f"(v5.y) : "l"(p));
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+1536];" : "=f"(v6.x), "=f"(v6.y) : "l"(p));
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+128];" : "=f"(v7.x), "=f"(v7.y) : "l"(p));
                a0 += v0.x + v0.y; a1 += v1.x + v1.y; a2 += v2.x + v2.y; a3 += v3.x + v3.y;
                a4 += v4.x + v4.y; a5 += v5.x + v5.y; a6 += v6.x + v6.y; a7 += v7.x + v7.y;
            } else if (OP == 9) {   // 4 MUFU + 4 SHFL: do XU and the shuffle network overlap?
                asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
                asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
                a4 = __shfl_sync(0xffffffffu, a4, (lane + 2) & 31); a5 = __shfl_sync(0xffffffffu, a5, (lane + 2) & 31);
                a6 = __shfl_sync(0xffffffffu, a6, (lane + 2) & 31); a7 = __shfl_sync(0xffffffffu, a7, (lane + 2) & 31);
            } else if (OP == 10) {   // 4 MUFU + 4 LDS.64
                asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
                asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
                const float2 *p = sm + lane + (i & 1);
                float2 v0, v1, v2, v3;
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v0.x), "=f"(v0.y) : "l"(p));
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+256];" : "=f"(v1.x), "=f"(v1.y) : "l"(p));
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+512];" : "=f"(v2.x), "=f"(v2.y) : "l"(p));
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+768];" : "=f"(v3.x), "=f"(v3.y) : "l"(p));
                a4 += v0.x; a5 += v1.y; a6 += v2.x; a7 += v3.y;
            } else if (OP == 11) {   // 8 LDS.64 only (4 FADD)
                const float2 *p = sm + lane + (i & 1);
                float2 v0, v1, v2, v3, v4, v5, v6, v7;
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v0.x), "=f"(v0.y) : "l"(p));
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+256];" : "=f"(v1.x), "=f"(v1.y) : "l"(p));
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+512];" : "=f"(v2.x), "=f"(v2.y) : "l"(p));
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+768];" : "=f"(v3.x), "=f"(v3.y) : "l"(p));
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+1024];" : "=f"(v4.x), "=f"(v4.y) : "l"(p));
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+1280];" : "=f"(v5.x), "=f"(v5.y) : "l"(p));
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+1536];" : "=f"(v6.x), "=f"(v6.y) : "l"(p));
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+128];" : "=f"(v7.x), "=f"(v7.y) : "l"(p));
                a0 += v0.x + v1.y; a1 += v2.x + v3.y; a2 += v4.x + v5.y; a3 += v6.x + v7.y;
            } else if (OP == 12) {   // 4 SHFL + 4 LDS.64
                a4 = __shfl_sync(0xffffffffu, a4, (lane + 2) & 31); a5 = __shfl_sync(0xffffffffu, a5, (lane + 2) & 31);
                a6 = __shfl_sync(0xffffffffu, a6, (lane + 2) & 31); a7 = __shfl_sync(0xffffffffu, a7, (lane + 2) & 31);
                const float2 *p = sm + lane + (i & 1);
                float2 v0, v1, v2, v3;
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v0.x), "=f"(v0.y) : "l"(p));
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+256];" : "=f"(v1.x), "=f"(v1.y) : "l"(p));
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+512];" : "=f"(v2.x), "=f"(v2.y) : "l"(p));
                asm volatile("ld.shared.v2.f32 {%0,%1}, [%2+768];" : "=f"(v3.x), "=f"(v3.y) : "l"(p));
                a0 += v0.x; a1 += v1.y; a2 += v2.x; a3 += v3.y;
            } else if (OP == 13) {   // 4 MUFU + 8 FFMA2 (pipes overlap?)
                asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
                asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
                float2 p2 = make_float2(a4, a5), p3 = make_float2(a6, a7);
                const float2 c = make_float2(0.999f, 0.999f), d = make_float2(1e-3f, 1e-3f);
                p2 = __ffma2_rn(p2, c, d); p3 = __ffma2_rn(p3, c, d); p2 = __ffma2_rn(p2, c, d); p3 = __ffma2_rn(p3, c, d);
                p2 = __ffma2_rn(p2, c, d); p3 = __ffma2_rn(p3, c, d); p2 = __ffma2_rn(p2, c, d); p3 = __ffma2_rn(p3, c, d);
                a4 = p2.x; a5 = p2.y; a6 = p3.x; a7 = p3.y;
            } else if (OP == 14) {   // FFMA2 with three distinct packed register operands (6 registers read per instruction)
                float2 p0 = make_float2(a0, a1), p1 = make_float2(a2, a3), p2 = make_float2(a4, a5), p3 = make_float2(a6, a7);
                p0 = __ffma2_rn(p1, p2, p0); p1 = __ffma2_rn(p2, p3, p1); p2 = __ffma2_rn(p3, p0, p2); p3 = __ffma2_rn(p0, p1, p3);
                p0 = __ffma2_rn(p1, p2, p0); p1 = __ffma2_rn(p2, p3, p1); p2 = __ffma2_rn(p3, p0, p2); p3 = __ffma2_rn(p0, p1, p3);
                a0 = p0.x; a1 = p0.y; a2 = p1.x; a3 = p1.y; a4 = p2.x; a5 = p2.y; a6 = p3.x; a7 = p3.y;
            } else if (OP == 15) {   // FMUL2 with two distinct packed register operands
                float2 p0 = make_float2(a0, a1), p1 = make_float2(a2, a3), p2 = make_float2(a4, a5), p3 = make_float2(a6, a7);
                p0 = __fmul2_rn(p1, p2); p1 = __fmul2_rn(p2, p3); p2 = __fmul2_rn(p3, p0); p3 = __fmul2_rn(p0, p1);
                p0 = __fmul2_rn(p1, p2); p1 = __fmul2_rn(p2, p3); p2 = __fmul2_rn(p3, p0); p3 = __fmul2_rn(p0, p1);
                a0 = p0.x; a1 = p0.y; a2 = p1.x; a3 = p1.y; a4 = p2.x; a5 = p2.y; a6 = p3.x; a7 = p3.y;
            } else if (OP == 8) {   // mix: the far-field chain (2 pairs): 11 packed + 4 MUFU
                float2 dx = make_float2(a0, a1), dy = make_float2(a2, a3), fx = make_float2(a4, a5), fy = make_float2(a6, a7);
                const float2 nc = make_float2(-18.f, -18.f), rc = make_float2(10.8f, 10.8f);
                float2 d2 = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
                float2 inv; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv.x) : "f"(d2.x)); asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv.y) : "f"(d2.y));
                float2 arg = __ffma2_rn(d2, __fmul2_rn(inv, nc), rc);
                float2 e; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(arg.x)); asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(arg.y));
                float2 c = __fmul2_rn(e, inv);
                fx = __ffma2_rn(c, dx, fx); fy = __ffma2_rn(c, dy, fy);
                dx = __fadd2_rn(dx, make_float2(1e-3f, 1e-3f)); dy = __fadd2_rn(dy, make_float2(1e-3f, 1e-3f));
                a0 = dx.x; a1 = dx.y; a2 = dy.x; a3 = dy.y; a4 = fx.x; a5 = fx.y; a6 = fy.x; a7 = fy.y;
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

template <int OP> void run(const char *name, double ops_per_iter, int threads, int blocks_per_sm)
{
    int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    const int blocks = p.multiProcessorCount * blocks_per_sm;
    float *buf; cudaMalloc(&buf, sizeof(float) * blocks * threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0); k<OP><<<blocks, threads>>>(buf, ITERS); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    const double warp_insts = ops_per_iter * 4.0 * ITERS * (threads / 32.0) * blocks_per_sm;   // per SM
    const double cycles = best * 1e-3 * clk * 1e3;
    printf("%-28s threads=%4d x%d/SM: %.3f ms, %.2f cycles per warp-instruction per SM  (%.2f per SMSP)\n", name, threads, blocks_per_sm, best,
           cycles / warp_insts, 4.0 * cycles / warp_insts);
    cudaFree(buf);
}

int main()
{
    for (int t : {128, 512}) {
        run<0>("MUFU.RSQ", 8, t, 2); run<1>("MUFU.EX2", 8, t, 2); run<2>("MUFU.RCP", 8, t, 2);
        run<3>("FFMA imm", 8, t, 2); run<4>("FFMA reg", 8, t, 2); run<5>("FFMA2 (per packed inst)", 8, t, 2);
        run<6>("SHFL.IDX", 8, t, 2); run<11>("LDS.64 x8 (+4 FADD)", 8, t, 2); run<8>("far chain (15 inst/2 pairs)", 15, t, 2);
        run<9>("4 MUFU + 4 SHFL (per 8)", 8, t, 2); run<10>("4 MUFU + 4 LDS.64 (per 8)", 8, t, 2); run<12>("4 SHFL + 4 LDS.64 (per 8)", 8, t, 2);
        run<13>("4 MUFU + 8 FFMA2 (per 12)", 12, t, 2);
        run<14>("FFMA2, 3 packed register operands", 8, t, 2); run<15>("FMUL2, 2 packed register operands", 8, t, 2);
    }
    return 0;
}
