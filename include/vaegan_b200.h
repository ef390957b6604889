/*
 * vaegan_b200 — C-ABI of the B200-native (sm_100a) VAE-GAN training-step kernels.
 *
 * This is the drop-in boundary for the hot path of the reference
 * (viniciusmenesessouza/VAE-GAN-based-model-for-image-generation-and-denoising):
 *   - main_vae.py:20-58   ConvBlock / Encoder        (Conv2d k4 s2 p0 + BatchNorm2d + LeakyReLU, two Linear heads)
 *   - gan_code.py:16-54   Generator (= Decoder)      (ConvTranspose2d + BatchNorm2d + ReLU, ConvT k3 + Tanh)
 *   - gan_code.py:56-89   Discriminator              (Conv2d k4 s2 p1 + BatchNorm2d + LeakyReLU(0.2), Conv + Sigmoid)
 *   - vaegan_code.py:74-135 reparameterisation, instance noise, BCE / MSE / KL losses, Adam
 * The reference reaches all of this arithmetic through torch (ATen / cuDNN / cuBLAS / oneDNN); every entry point
 * below names the torch call site it replaces.
 *
 * Conventions
 *   - plain C: raw device pointers, sizes, POD structs; no C++/torch types.
 *   - the caller owns every buffer (inputs, outputs, workspaces); the library never allocates device memory.
 *   - all calls are asynchronous on `stream` (a cudaStream_t passed as void*), never synchronise, and are
 *     CUDA-graph capturable.
 *   - return value: VG_OK (0) or a negative VG_ERR_* code; vg_last_error() returns a thread-local message.
 *   - there is NO CPU fallback: on a device that is not sm_100 the compute entry points return VG_ERR_ARCH.
 *   - activations are NHWC ("channels-last") between kernels, dtype = VgDType; fp32 statistics / losses / optimizer.
 *
 * Convolution geometry.  Conv2d and ConvTranspose2d of the reference are the same three contractions over a
 * "big" (high-resolution) and a "small" (low-resolution) tensor that share ONE weight tensor
 * w[small_c][big_c][k][k]  (Conv2d.weight = [Cout,Cin,k,k] with out=small; ConvTranspose2d.weight = [Cin,Cout,k,k]
 * with in=small):
 *      down : small[b,oy,ox,sc] = sum  big[b, oy*s-p+ky, ox*s-p+kx, bc] * w[sc][bc][ky][kx]
 *      up   : big[b,y,x,bc]     = sum  small[b,(y+p-ky)/s,(x+p-kx)/s, sc] * w[sc][bc][ky][kx]   (divisible taps only)
 *      wgrad: dw[sc][bc][ky][kx] += sum small[b,oy,ox,sc] * big[b, oy*s-p+ky, ox*s-p+kx, bc]
 *   Conv2d          : forward = down, input-gradient = up,   weight-gradient = wgrad(small = dY, big = X)
 *   ConvTranspose2d : forward = up,   input-gradient = down, weight-gradient = wgrad(small = X,  big = dY)
 */
#ifndef VAEGAN_B200_H
#define VAEGAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VG_OK 0
#define VG_ERR_SHAPE (-1)     /* unsupported / inconsistent geometry */
#define VG_ERR_ALIGN (-2)     /* pointer or extent violates an alignment requirement */
#define VG_ERR_ARCH (-3)      /* device is not sm_100 (no fallback) */
#define VG_ERR_WORKSPACE (-4) /* workspace too small */
#define VG_ERR_CUDA (-5)      /* a CUDA runtime / driver call failed */
#define VG_ERR_ARG (-6)       /* bad enum / null pointer */

typedef enum VgDType { VG_F32 = 0, VG_BF16 = 1 } VgDType;
typedef enum VgAct { VG_ACT_NONE = 0, VG_ACT_RELU = 1, VG_ACT_LEAKY = 2, VG_ACT_TANH = 3, VG_ACT_SIGMOID = 4 } VgAct;

typedef struct VgConvGeom {
    int32_t batch;
    int32_t big_h, big_w, big_c;       /* Conv2d input  / ConvTranspose2d output */
    int32_t small_h, small_w, small_c; /* Conv2d output / ConvTranspose2d input  */
    int32_t kernel, stride, pad;
} VgConvGeom;

const char* vg_last_error(void);
int vg_version(void);
/* 0 when the current device is sm_100 and the kernels can run, VG_ERR_ARCH / VG_ERR_CUDA otherwise. */
int vg_device_check(void);

/* ---- weight packing (bf16 tensor-core path) ---------------------------------------------------------------
 * fp32 master w[small_c][big_c][k][k]  ->  wd[tap][small_c][big_c]  (operand of `down`)
 *                                      and wu[tap][big_c][small_c]  (operand of `up`),  tap = ky*k+kx, bf16.
 * Either output may be NULL.  Replaces the implicit weight reads of nn.Conv2d / nn.ConvTranspose2d
 * (main_vae.py:23, gan_code.py:21-49, 61-84). */
int vg_pack_weights_bf16(const VgConvGeom* g, const float* w, void* wd, void* wu, void* stream);

/* ---- convolution contractions -------------------------------------------------------------------------------
 * dtype selects the arithmetic path: VG_BF16 = TMA-fed tcgen05/TMEM implicit GEMM (bf16 operands, fp32
 * accumulate) with packed weights `w` (wd for down, wu for up); VG_F32 = fp32 CUDA-core implicit GEMM reading
 * the fp32 master weights directly.  `bias` (fp32[out channels]) may be NULL.  `out_f32` != 0 writes the
 * result as fp32 regardless of dtype (used for mu / logvar).
 * down  replaces F.conv2d            (main_vae.py:28, gan_code.py:89) and the dgrad of F.conv_transpose2d
 * up    replaces F.conv_transpose2d  (gan_code.py:54)               and the dgrad of F.conv2d
 * wgrad replaces the weight gradient of both (autograd of vaegan_code.py:104,133) */
int vg_conv_down(const VgConvGeom* g, VgDType dtype, const void* big, const void* w, const float* bias, void* small,
                 int out_f32, void* stream);
int vg_conv_up(const VgConvGeom* g, VgDType dtype, const void* small, const void* w, void* big, void* stream);
/* dw (fp32, reference layout [small_c][big_c][k][k]) is ACCUMULATED into (+=); zero it for a fresh gradient. */
int vg_conv_wgrad(const VgConvGeom* g, VgDType dtype, const void* small, const void* big, float* dw, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VAEGAN_B200_H */
