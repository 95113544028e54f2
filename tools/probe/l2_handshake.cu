// l2_handshake.cu -- what a block-to-block hand-over through L2 costs on B200 (design input of resident.cuh).
//
//   1. load latency of data another SM wrote (dependent chain): weak ld.global, ld.global.cg, ld.relaxed.gpu
//   2. flag ping-pong between block 0 and block b (one thread each): relaxed store / relaxed poll, with and without
//      release / acquire fences, and "flag in the data" (the value itself is polled)
//   3. store -> fence -> flag store: cost of fence.acq_rel.gpu after N outstanding stores
// Build: nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o l2_handshake l2_handshake.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned ld_relaxed(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned *p, unsigned v)
{
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- 1. dependent loads of an index chain written by `writer` block, read by `reader` block --------------
template <int KIND>
__global__ void chase(unsigned *buf, int n, int writer, int reader, unsigned *flag, long long *out)
{
    if ((int)blockIdx.x == writer && threadIdx.x == 0) {
        for (int i = 0; i < n; i++) buf[i * 32] = (unsigned)(((i + 1) % n) * 32);      // one 128 B line per hop
        __threadfence();
        st_relaxed(flag, 1u);
    }
    if ((int)blockIdx.x == reader && threadIdx.x == 0) {
        while (ld_relaxed(flag) == 0) {}
        __threadfence();
        unsigned idx = 0;
        const long long t0 = clock64();
        for (int i = 0; i < n; i++) {
            if (KIND == 1) asm volatile("ld.global.u32 %0, [%1];" : "=r"(idx) : "l"(buf + idx) : "memory");
            if (KIND == 2) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(idx) : "l"(buf + idx) : "memory");
            if (KIND == 3) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(idx) : "l"(buf + idx) : "memory");
            if (KIND == 4) asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(idx) : "l"(buf + idx) : "memory");
        }
        const long long t1 = clock64();
        out[0] = t1 - t0;
        out[1] = idx;
    }
}

// ---- 2. ping-pong ------------------------------------------------------------------------------------------
// MODE 0: relaxed store / relaxed poll.  MODE 1: st.release / poll + fence.acq_rel.  MODE 2: as 1, and 64 data stores
// by the same thread before every release (what a block's update leaves outstanding).
template <int MODE>
__global__ void pingpong(unsigned *fa, unsigned *fb, unsigned *data, int other, int rounds, long long *out)
{
    if (threadIdx.x != 0) return;
    const bool a = blockIdx.x == 0, b = (int)blockIdx.x == other;
    if (!a && !b) return;
    unsigned *mine = a ? fa : fb, *theirs = a ? fb : fa;
    unsigned *d = data + (a ? 0 : 1 << 16);
    const long long t0 = clock64();
    for (int r = 1; r <= rounds; r++) {
        if (b) { while (ld_relaxed(theirs) < (unsigned)r) {} if (MODE) asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
        if (MODE == 2)
            for (int i = 0; i < 64; i++) d[i * 32] = r;
        if (MODE) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(mine), "r"((unsigned)r) : "memory");
        else st_relaxed(mine, (unsigned)r);
        if (a) { while (ld_relaxed(theirs) < (unsigned)r) {} if (MODE) asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
    }
    if (a) out[0] = clock64() - t0;
}

int main()
{
    unsigned *buf, *flags, *data;
    long long *out, h[2];
    cudaMalloc(&buf, 1 << 22);
    cudaMalloc(&flags, 4096);
    cudaMalloc(&data, 1 << 20);
    cudaMalloc(&out, 64);
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int n = 512;
    const int others[] = {1, 2, nsm / 4, nsm / 2, nsm / 2 + 1, nsm - 1};
    const char *names[] = {"", "ld.global (weak, L1)", "ld.global.cg", "ld.relaxed.gpu", "ld.global.L1::no_allocate"};
    for (int w : others) {
        for (int kind = 1; kind <= 4; kind++) {
            cudaMemset(flags, 0, 4096);
            void (*k)(unsigned *, int, int, int, unsigned *, long long *) =
                kind == 1 ? chase<1> : kind == 2 ? chase<2> : kind == 3 ? chase<3> : chase<4>;
            k<<<nsm, 32>>>(buf, n, w, 0, flags, out);
            cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            printf("chase  writer block %3d -> reader block 0  %-28s %7.1f cycles / load\n", w, names[kind], (double)h[0] / n);
        }
    }
    for (int o : others) {
        for (int mode = 0; mode < 3; mode++) {
            cudaMemset(flags, 0, 4096);
            const int rounds = 2000;
            void (*k)(unsigned *, unsigned *, unsigned *, int, int, long long *) = mode == 0 ? pingpong<0> : mode == 1 ? pingpong<1> : pingpong<2>;
            k<<<nsm, 32>>>(flags, flags + 64, data, o, rounds, out);
            cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
            printf("pingpong block 0 <-> block %3d  mode %d (%s)  %7.1f cycles one way\n", o, mode,
                   mode == 0 ? "relaxed" : mode == 1 ? "release/acquire" : "64 stores + release/acquire", (double)h[0] / rounds / 2);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
