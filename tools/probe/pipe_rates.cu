// Probe: issue rates of the instructions the median / homogeneity kernels are bound by, on the GPU at hand.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
// Each kernel runs ILP independent dependency chains per thread, 1024 threads per SM-resident block set.
#include <cstdio>
#include <cuda_runtime.h>

#define ILP 8
#define ITERS 4096

template <int OP>
__global__ void __launch_bounds__(256) k(float* out, float seed, int iseed) {
    float f[ILP];
    int v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { f[i] = seed + threadIdx.x * 0.001f + i; v[i] = iseed + threadIdx.x * 7 + i * 13; }
    float g = seed * 0.5f, h = seed * 0.25f;
    int w = iseed * 3, z = iseed * 5, one = iseed / 3, mone = -one;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) { if (i & 1) asm volatile("min.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g)); else asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(h)); if (it & 1) { float t = f[i]; f[i] = f[(i + 1) % ILP]; f[(i + 1) % ILP] = t; } }                                   // FMNMX
            if (OP == 1) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(g), "f"(h));                         // FMNMX3
            if (OP == 2) { if (i & 1) asm volatile("min.s32 %0, %0, %1;" : "+r"(v[i]) : "r"(w)); else asm volatile("max.s32 %0, %0, %1;" : "+r"(v[i]) : "r"(z)); if (it & 1) { int t = v[i]; v[i] = v[(i + 1) % ILP]; v[(i + 1) % ILP] = t; } }                                     // VIMNMX
            if (OP == 3) { asm volatile("min.s32 %0, %0, %1;" : "+r"(v[i]) : "r"(w)); asm volatile("min.s32 %0, %0, %1;" : "+r"(v[i]) : "r"(z)); }                             // VIMNMX3
            if (OP == 4) asm volatile("mad.lo.s32 %0, %0, %2, %1;" : "+r"(v[i]) : "r"(w), "r"(one));       // IMAD (x*one+w)
            if (OP == 5) asm volatile("add.s32 %0, %0, %1;" : "+r"(v[i]) : "r"(w));             // IADD3
            if (OP == 6) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g));                               // FADD
            if (OP == 7) { asm volatile("min.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g)); asm volatile("mad.lo.s32 %0, %0, %2, %1;" : "+r"(v[i]) : "r"(w), "r"(one)); }   // FMNMX + IMAD
            if (OP == 8) { asm volatile("min.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(h)); }    // FMNMX + FADD
            if (OP == 9) { asm volatile("min.s32 %0, %0, %1;" : "+r"(v[i]) : "r"(w)); asm volatile("mad.lo.s32 %0, %0, %2, %1;" : "+r"(v[i]) : "r"(z), "r"(one));
                           asm volatile("mad.lo.s32 %0, %0, %2, %1;" : "+r"(v[i]) : "r"(w), "r"(mone)); }  // VIMNMX + 2 IMAD
            if (OP == 10) asm volatile("{ .reg .pred p; setp.le.f32 p, %0, %1; selp.f32 %0, %0, %2, p; }" : "+f"(f[i]) : "f"(g), "f"(h));  // FSETP+FSEL
            if (OP == 11) asm volatile("prmt.b32 %0, %0, %1, 0x7610;" : "+r"(v[i]) : "r"(w));    // PRMT
            if (OP == 12) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(g));                              // FMUL (3-reg)
            if (OP == 13) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(g), "f"(h));
            if (OP == 14) asm volatile("dp2a.lo.u32.u32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(w), "r"(z));                 // IDP.2A
            if (OP == 15) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(w), "r"(z));                    // IDP.4A
            if (OP == 16) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[i]) : "r"(w), "r"(z));                  // LOP3
            if (OP == 17) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(w), "r"(one));                // SHF
            if (OP == 18) { double t = __longlong_as_double(((long long)v[i] << 32) | 0x3ff0000000000000LL); asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(t) : "d"((double)g), "d"((double)h)); v[i] = (int)(__double_as_longlong(t) >> 32); }   // DFMA (+moves)
            if (OP == 19) { float t; asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(t) : "r"(v[i])); f[i] += t; v[i] ^= __float_as_int(f[i]); }     // I2F + FADD + LOP
            if (OP == 20) { double t; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(t) : "f"(f[i])); asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[i]) : "d"(t)); }   // F2F x2
            if (OP == 21) asm volatile("mad.lo.s32 %0, %0, 3, %1;" : "+r"(v[i]) : "r"(w));                               // IMAD imm
            if (OP == 22) asm volatile("{ .reg .pred p; setp.le.f32 p, %0, %1; @p add.rn.f32 %0, %0, %2; }" : "+f"(f[i]) : "f"(g), "f"(h));   // FSETP + predicated FADD
            if (OP == 23) asm volatile("add.rn.f32 %0, %0, 0f3F800000;" : "+f"(f[i]));                                   // FADD imm                                // FFMA (3-reg)
        }
    }
    float s = 0; int t = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { s += f[i]; t += v[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + t;
}

template <int OP>
void run(const char* name, int ops_per_iter, float* d) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<OP><<<sms * 4, 256>>>(d, 1.0f, 3);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<OP><<<sms * 4, 256>>>(d, 1.0f, 3);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double warp_inst = (double)sms * 4 * 8 * ITERS * ILP * ops_per_iter;      // warps x iterations x ops
    double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-22s %8.3f ms   %.3f warp-inst/clk/SMSP (nominal clock %d MHz)\n", name, ms, warp_inst / cycles / (sms * 4), clk / 1000);
}

int main() {
    float* d; cudaMalloc(&d, 148 * 4 * 256 * 4 * 2);
    run<0>("FMNMX", 1, d); run<1>("FMNMX3", 1, d); run<2>("VIMNMX", 1, d); run<3>("2xVIMNMX(->3?)", 2, d);
    run<4>("IMAD", 1, d); run<5>("IADD3", 1, d); run<6>("FADD", 1, d); run<12>("FMUL", 1, d); run<13>("FFMA", 1, d);
    run<7>("FMNMX+IMAD", 2, d); run<8>("FMNMX+FADD", 2, d); run<9>("VIMNMX+2IMAD", 3, d); run<10>("FSETP+FSEL", 2, d); run<11>("PRMT", 1, d);
    run<14>("IDP.2A", 1, d); run<15>("IDP.4A", 1, d); run<16>("LOP3", 1, d); run<17>("SHF", 1, d); run<18>("DFMA(+2 mov)", 1, d);
    run<19>("I2F+FADD+LOP", 3, d); run<20>("F2F.64.32+F2F.32.64", 2, d); run<21>("IMAD imm", 1, d); run<22>("FSETP+@FADD", 2, d); run<23>("FADD imm", 1, d);
    return 0;
}
