// Differential check: wide_node_hits() on the device vs the host emulation, random nodes and rays.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../build-up-phase_b200/csrc/wide_bvh.cuh"
using namespace rt;
struct Case { WNode n; float o[3], d[3]; float absmax[3]; float tbest; };
__global__ void k(const Case* c, uint32_t* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
    RayBox rb; raybox_setup(rb, c[i].o, c[i].d, c[i].absmax[0], c[i].absmax[1], c[i].absmax[2]);
    const uint4* w = reinterpret_cast<const uint4*>(&c[i].n);
    out[i] = wide_node_hits(rb, w[0], w[1], w[2], w[3], w[4], 0.0f, c[i].tbest);
}
static uint32_t s = 12345; static uint32_t rnd() { s = s * 1664525u + 1013904223u; return s >> 8; }
static float frand(float a, float b) { return a + (b - a) * (rnd() & 0xFFFF) / 65535.0f; }
int main() {
    const int N = 200000;
    std::vector<Case> cs(N);
    for (auto& c : cs) {
        c.n.px = frand(-5, 5); c.n.py = frand(-5, 5); c.n.pz = frand(-5, 5);
        c.n.ex = 110 + rnd() % 12; c.n.ey = 110 + rnd() % 12; c.n.ez = 110 + rnd() % 12; c.n.imask = 0;
        c.n.child_base = rnd(); c.n.prim_base = rnd();
        int off = 0;
        for (int s2 = 0; s2 < 8; ++s2) {
            int kind = rnd() % 4;
            if (kind == 0) c.n.meta[s2] = 0;
            else if (kind == 1) { c.n.meta[s2] = 0x20 | (24 + s2); c.n.imask |= 1u << s2; }
            else { int cnt = 1 + rnd() % 3; c.n.meta[s2] = (((1u << cnt) - 1) << 5) | off; off += cnt; }
            uint8_t* q[6] = {c.n.qlox, c.n.qloy, c.n.qloz, c.n.qhix, c.n.qhiy, c.n.qhiz};
            for (int k2 = 0; k2 < 3; ++k2) { int a = rnd() % 256, b = rnd() % 256; if (a > b) { int t = a; a = b; b = t; } q[k2][s2] = a; q[3 + k2][s2] = b; }
        }
        for (int k2 = 0; k2 < 3; ++k2) { c.o[k2] = frand(-8, 8); c.d[k2] = frand(-1, 1); c.absmax[k2] = 8; }
        if (rnd() % 8 == 0) c.d[rnd() % 3] = 0.0f;
        c.tbest = frand(0.5f, 100.0f);
    }
    Case* dc; uint32_t* dout; cudaMalloc(&dc, sizeof(Case) * N); cudaMalloc(&dout, 4 * N);
    cudaMemcpy(dc, cs.data(), sizeof(Case) * N, cudaMemcpyHostToDevice);
    k<<<(N + 127) / 128, 128>>>(dc, dout, N);
    std::vector<uint32_t> out(N); cudaMemcpy(out.data(), dout, 4 * N, cudaMemcpyDeviceToHost);
    printf("cuda: %s\n", cudaGetErrorString(cudaGetLastError()));
    int bad = 0, nonzero = 0;
    for (int i = 0; i < N; ++i) {
        RayBox rb; raybox_setup(rb, cs[i].o, cs[i].d, 8, 8, 8);
        const uint4* w = reinterpret_cast<const uint4*>(&cs[i].n);
        uint32_t h = wide_node_hits(rb, w[0], w[1], w[2], w[3], w[4], 0.0f, cs[i].tbest);
        if (h) ++nonzero;
        if (h != out[i]) { if (bad < 10) printf("case %d host %08x dev %08x oct %u\n", i, h, out[i], rb.oct); ++bad; }
    }
    printf("cases %d nonzero %d mismatches %d\n", N, nonzero, bad);
    return 0;
}
