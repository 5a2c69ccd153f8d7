// Training-image augmentation of the reference's dataset (dataset.py:66-79) for a whole batch on the device:
//     image = cv2.imread(...).astype(np.float32) / 255.              -> here: the decoded u8 HWC BGR image
//     imgproc.random_rotate(image, [0, 90, 180, 270])                  (imgproc.py:1937-1963: cv2.warpAffine about (w//2, h//2))
//     imgproc.random_horizontally_flip / random_vertically_flip        (imgproc.py:1966-2001: cv2.flip)
//     cv2.cvtColor(BGR2RGB); imgproc.image_to_tensor(.., False, False) (imgproc.py:1540-1567: HWC -> CHW)
// Rotations by multiples of 90 degrees are exact pixel copies in cv2 (its fixed-point coordinates snap to integers), so the
// whole chain is ONE gather: dst[c][y][x] = src[sy][sx][2 - c] / 255 with an integer (sx, sy) per (angle, flips); pixels that
// rotate in from outside the canvas are 0 (BORDER_CONSTANT). Note the reference's centre (w // 2, h // 2): the rotated
// image is shifted by one pixel for even sizes, and this kernel reproduces that. Bit-exact (oracle/augment.py,
// tests/golden/augment.npz from the reference's own functions).
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/resr.h"
#include "errors.h"

namespace resr {

__global__ void __launch_bounds__(256) augment_u8_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, const int* __restrict__ ops,
                                                         int H, int W) {
    const int b = blockIdx.z;
    const int op = ops[b];
    const int ai = op & 3, hf = (op >> 2) & 1, vf = (op >> 3) & 1;
    const int cx = W / 2, cy = H / 2;
    const size_t HW = static_cast<size_t>(H) * W;
    const uint8_t* im = src + static_cast<size_t>(b) * HW * 3;
    float* out = dst + static_cast<size_t>(b) * HW * 3;
    for (int y = blockIdx.y; y < H; y += gridDim.y) {
        for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < W; x += gridDim.x * blockDim.x) {
            const int xf = hf ? W - 1 - x : x, yf = vf ? H - 1 - y : y;   // cv2.flip is applied AFTER the rotation
            int sx, sy;
            if (ai == 0) { sx = xf; sy = yf; }
            else if (ai == 1) { sx = cx + cy - yf; sy = xf - cx + cy; }      // 90 degrees (counter-clockwise, cv2 convention)
            else if (ai == 2) { sx = 2 * cx - xf; sy = 2 * cy - yf; }
            else { sx = yf + cx - cy; sy = cx + cy - xf; }
            float r = 0.f, g = 0.f, bl = 0.f;
            if (sx >= 0 && sx < W && sy >= 0 && sy < H) {
                const uint8_t* p = im + (static_cast<size_t>(sy) * W + sx) * 3;
                bl = __fdiv_rn(static_cast<float>(p[0]), 255.f);
                g = __fdiv_rn(static_cast<float>(p[1]), 255.f);
                r = __fdiv_rn(static_cast<float>(p[2]), 255.f);
            }
            const size_t o = static_cast<size_t>(y) * W + x;
            out[o] = r; out[o + HW] = g; out[o + 2 * HW] = bl;
        }
    }
}

}  // namespace resr

extern "C" int resr_augment_batch_u8(const unsigned char* images_bgr_hwc, float* out_rgb_nchw, const int* ops, int b, int h, int w,
                                     void* stream) {
    using namespace resr;
    if (!images_bgr_hwc || !out_rgb_nchw || !ops) return set_error(RESR_E_INVALID, "null argument");
    if (b <= 0 || h <= 0 || w <= 0 || b > 65535) return set_error(RESR_E_INVALID, "bad shape");
    const dim3 grid((w + 255) / 256, h < 1024 ? h : 1024, b);
    augment_u8_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(images_bgr_hwc, out_rgb_nchw, ops, h, w);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(RESR_E_CUDA, "augment launch: %s", cudaGetErrorString(e));
    return RESR_OK;
}
