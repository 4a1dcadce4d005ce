#include <cstdio>
#include "bsw_kernels.cuh"
using namespace bswk;
__global__ void k(KParams P, int *out) {
    int q = threadIdx.x + 1;
    out[q] = pair_band(P, q);
}
int main() {
    int *d; cudaMalloc(&d, 1024 * 4);
    int h[1024];
    KParams P{6, 1, 6, 1, 100, 5, 1, 4, -1, 3, 1};
    k<<<1, 64>>>(P, d); cudaMemcpy(h, d, 4096, cudaMemcpyDeviceToHost);
    printf("w=3 default: "); for (int q = 1; q <= 8; ++q) printf("%d ", h[q]); printf("\n");
    KParams P2{6, 2, 6, 2, 30, 5, 1, 4, -1, 100, 1};
    k<<<1, 64>>>(P2, d); cudaMemcpy(h, d, 4096, cudaMemcpyDeviceToHost);
    printf("gape2 w=100: "); for (int q = 50; q <= 56; ++q) printf("%d ", h[q]); printf("\n");
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
}
