// TMA probe: which (box width, start coordinate) combinations does cp.async.bulk.tensor.2d accept for a u8 tensor?
// usage: tma_probe <boxw> <xstart> <ystart>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
struct Map { CUtensorMap m; };
__global__ void k(const __grid_constant__ Map map, int boxw, int boxh, int xs, int ys, unsigned char* out, int* flag) {
    extern __shared__ __align__(128) unsigned char tile[];
    __shared__ __align__(8) unsigned long long mbar;
    const unsigned mb = (unsigned)__cvta_generic_to_shared(&mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(mb));
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned dst = (unsigned)__cvta_generic_to_shared(tile);
        const unsigned long long desc = reinterpret_cast<unsigned long long>(&map.m);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(mb), "r"(boxw * boxh) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                     :: "r"(dst), "l"(desc), "r"(xs), "r"(ys), "r"(mb) : "memory");
    }
    unsigned done = 0;
    for (int spin = 0; spin < (1 << 22) && !done; ++spin)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(mb) : "memory");
    if (!done) { if (threadIdx.x == 0) *flag = 1; return; }
    for (int i = threadIdx.x; i < boxw * boxh; i += blockDim.x) out[i] = tile[i];
}
int main(int argc, char** argv) {
    const int boxw = atoi(argv[1]), xs = atoi(argv[2]), ys = atoi(argv[3]);
    const int W = 640, H = 64, boxh = 8;
    std::vector<unsigned char> img(W * H);
    for (int i = 0; i < W * H; ++i) img[i] = (unsigned char)((i * 7 + i / W) & 0xff);
    unsigned char *d, *o; int* f;
    cudaMalloc(&d, W * H); cudaMalloc(&o, 65536); cudaMalloc(&f, 4); cudaMemset(f, 0, 4);
    cudaMemcpy(d, img.data(), W * H, cudaMemcpyHostToDevice);
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    Map map;
    const cuuint64_t gdim[2] = {(cuuint64_t)W, (cuuint64_t)H};
    const cuuint64_t gstr[1] = {(cuuint64_t)W};
    const cuuint32_t box[2] = {(cuuint32_t)boxw, (cuuint32_t)boxh};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = ((encode_fn)fn)(&map.m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("boxw %d xs %d ys %d: encode failed %d\n", boxw, xs, ys, (int)r); return 0; }
    k<<<1, 128, boxw * boxh>>>(map, boxw, boxh, xs, ys, o, f);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("boxw %d xs %d ys %d: kernel error: %s\n", boxw, xs, ys, cudaGetErrorString(e)); return 0; }
    int hf; cudaMemcpy(&hf, f, 4, cudaMemcpyDeviceToHost);
    std::vector<unsigned char> out(boxw * boxh);
    cudaMemcpy(out.data(), o, boxw * boxh, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < boxh; ++y) for (int x = 0; x < boxw; ++x) {
        const int gx = xs + x, gy = ys + y;
        const unsigned char want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? img[gy * W + gx] : 0;
        bad += out[y * boxw + x] != want;
    }
    printf("boxw %d xs %d ys %d: timeout %d mismatches %d\n", boxw, xs, ys, hf, bad);
    return 0;
}
