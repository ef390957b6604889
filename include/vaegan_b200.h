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
    /* Channel padding of the big side: the big TENSOR has big_c channels per pixel of which the first big_c_valid
     * exist in the weight tensor w[small_c][big_c_valid][k][k] (0 means big_c).  The bf16 path stores the 3-channel
     * image tensors with 16 channels (zeros above 3) so that they are legal TMA / UMMA operands. */
    int32_t big_c_valid;
} VgConvGeom;

const char* vg_last_error(void);
int vg_version(void);
/* 0 when the current device is sm_100 and the kernels can run, VG_ERR_ARCH / VG_ERR_CUDA otherwise. */
int vg_device_check(void);
/* kernels launched by this library in this process so far (host-side counter). */
long long vg_launch_count(void);

/* ---- weight packing (bf16 tensor-core path) ---------------------------------------------------------------
 * fp32 master w[small_c][big_c_valid][k][k]  ->  wd[tap][small_c][big_c]  (operand of `down`)
 *                                            and wu[tap][big_c][small_c]  (operand of `up`),  tap = ky*k+kx, bf16,
 * zero-filled for the padded channels big_c_valid <= c < big_c.
 * Either output may be NULL.  Replaces the implicit weight reads of nn.Conv2d / nn.ConvTranspose2d
 * (main_vae.py:23, gan_code.py:21-49, 61-84). */
int vg_pack_weights_bf16(const VgConvGeom* g, const float* w, void* wd, void* wu, void* stream);
/* The same for many layers in ONE launch (items is a host array; used after every optimizer step). */
typedef struct VgPackItem {
    const float* w;
    void* wd;
    void* wu;
    int32_t small_c, big_c, big_c_valid, kk; /* kk = kernel * kernel */
} VgPackItem;
int vg_pack_weights_multi(const VgPackItem* items, int n_items, void* stream);

/* ---- convolution contractions -------------------------------------------------------------------------------
 * dtype selects the arithmetic path: VG_BF16 = TMA-fed tcgen05/TMEM implicit GEMM (bf16 operands, fp32
 * accumulate) with packed weights `w` (wd for down, wu for up); VG_F32 = fp32 CUDA-core implicit GEMM reading
 * the fp32 master weights directly.  `bias` (fp32[out channels]) may be NULL.  `out_f32` != 0 writes the
 * result as fp32 regardless of dtype (used for mu / logvar).
 * down  replaces F.conv2d            (main_vae.py:28, gan_code.py:89) and the dgrad of F.conv_transpose2d
 * up    replaces F.conv_transpose2d  (gan_code.py:54)               and the dgrad of F.conv2d
 * wgrad replaces the weight gradient of both (autograd of vaegan_code.py:104,133) */
/* Optional scratch for vg_conv_down (ws may be NULL): with a buffer of vg_conv_down_workspace_bytes() the tensor-core
 * path splits a long reduction over many CTAs when the output has only a few tiles (e.g. the dgrad of the
 * generator's first ConvTranspose2d: batch x 16384 -> batch x nz). */
size_t vg_conv_down_workspace_bytes(const VgConvGeom* g);
int vg_conv_down(const VgConvGeom* g, VgDType dtype, const void* big, const void* w, const float* bias, void* small,
                 int out_f32, void* ws, size_t ws_bytes, void* stream);
int vg_conv_up(const VgConvGeom* g, VgDType dtype, const void* small, const void* w, void* big, void* stream);
/* ---- fused epilogues of the tensor-core contractions (bf16 path) ----------------------------------------------
 * The BatchNorm that follows a convolution in every (Conv|ConvT, BatchNorm2d, ReLU|LeakyReLU) triple of the
 * reference (main_vae.py:27-31, gan_code.py:19-51, 59-86) needs per-channel sums of the convolution output, and its
 * backward needs per-channel sums over the gradient the NEXT layer's dgrad produces.  Both ride the epilogue of the
 * producing contraction, so the activation is not re-read from HBM for the reduction:
 *   VG_EPI_BN_STATS  out = conv;  sums[g][0][c] += out, sums[g][1][c] += out^2            (forward)
 *   VG_EPI_BN_BWD    acc = dy;    out = dz = dy*act'(x*scale+shift);  sums[g][0][c] += dz,
 *                                 sums[g][1][c] += dz*(x-mean)*rstd                        (dgrad of the next layer)
 *   VG_EPI_ACT_BWD   out = dy*act'(x)                                                      (layer without BatchNorm)
 *   VG_EPI_ACT_FWD   out = act(conv)  (ReLU / LeakyReLU of a layer without BatchNorm; its backward takes act' from
 *                                      the sign of the activated output, so the raw output is never stored)
 *   VG_EPI_AFFINE_ACT_FWD  out = act(conv*scale[c] + shift[c])  (eval-mode BatchNorm + activation of the generation /
 *                                      validation passes, main_vae.py:348-374, vaegan_code.py:147-171: scale / shift =
 *                                      rows 2 / 3 of `stats` as vg_bn_eval_coeffs writes them; no separate pass)
 * `groups` = independent sub-batches along the batch axis with separate statistics; `sums` = fp32
 * [groups][2][channels], zero-initialised by the caller, accumulated with atomics; `x` = the saved raw convolution
 * output (same NHWC shape as this call's output); `stats` = [groups][4][channels] (mean, rstd, scale, shift).
 * vg_conv_epilogue_supported() tells whether a geometry takes the fused form (1) or the caller has to use the
 * stand-alone reduction kernels below (0). */
typedef enum VgEpilogueMode {
    VG_EPI_NONE = 0, VG_EPI_BN_STATS = 1, VG_EPI_BN_BWD = 2, VG_EPI_ACT_BWD = 3, VG_EPI_ACT_FWD = 4,
    VG_EPI_AFFINE_ACT_FWD = 5
} VgEpilogueMode;
typedef struct VgEpilogue {
    int32_t mode;      /* VgEpilogueMode */
    int32_t groups;
    int32_t channels;
    int32_t act;       /* VgAct of the layer whose backward is fused (modes 2, 3) */
    float slope;
    float* sums;
    const void* x;
    const float* stats;
} VgEpilogue;
int vg_conv_epilogue_supported(const VgConvGeom* g, VgDType dtype, int up, const VgEpilogue* ep);
int vg_conv_down_ex(const VgConvGeom* g, VgDType dtype, const void* big, const void* w, const float* bias, void* small,
                    const VgEpilogue* ep, void* stream);
int vg_conv_up_ex(const VgConvGeom* g, VgDType dtype, const void* small, const void* w, void* big,
                  const VgEpilogue* ep, void* stream);
/* dw (fp32, reference layout [small_c][big_c_valid][k][k]) is ACCUMULATED into (+=); zero it for a fresh gradient
 * (what autograd's `.grad +=` / loss.backward() after optimizer.zero_grad() of vaegan_code.py:103-104,131-133 does).
 * Layers with few weight tiles and a long pixel reduction split the reduction across SMs; the splits add their
 * tiles into dw with 16-byte vector reductions (fp32 atomics: the summation order, hence the last bits, vary from
 * run to run).  vg_conv_wgrad_workspace_bytes() reports the scratch the non-atomic variant (partial tiles + a
 * reduction kernel; experiment switch VG_WGRAD_ATOMIC=0) needs - 0 otherwise; ws may be NULL.
 * vg_conv_wgrad_ex flags: VG_WGRAD_OVERWRITE - dw = gradient; dw need not be initialised (a tile's single owner skips
 * the read-modify-write, split layers zero dw first).  VG_WGRAD_DST_ZERO - the caller guarantees dw is all zeros
 * (first gradient after optimizer.zero_grad()): same result as accumulating, without reading dw where one CTA owns
 * a tile and without the extra memset of the split layers. */
#define VG_WGRAD_OVERWRITE 1
#define VG_WGRAD_DST_ZERO 2
size_t vg_conv_wgrad_workspace_bytes(const VgConvGeom* g, VgDType dtype);
int vg_conv_wgrad(const VgConvGeom* g, VgDType dtype, const void* small, const void* big, float* dw, void* ws,
                  size_t ws_bytes, void* stream);
int vg_conv_wgrad_ex(const VgConvGeom* g, VgDType dtype, const void* small, const void* big, float* dw, void* ws,
                     size_t ws_bytes, int flags, void* stream);


/* ---- BatchNorm2d (training statistics), activations, layout edges ---------------------------------------------
 * All tensors NHWC viewed as [rows = B*H*W][C].  Reductions are two-stage and deterministic; the caller passes a
 * scratch buffer of vg_reduce_workspace_bytes() / vg_bn_bwd_workspace_bytes().
 * vg_bn_train_fwd replaces F.batch_norm(training=True) of nn.BatchNorm2d (main_vae.py:24,29; gan_code.py:22-46,
 * 65-81): batch mean / biased variance -> mean, rstd, and the fused affine  scale = gamma*rstd,
 * shift = beta - mean*scale;  running_mean / running_var (unbiased) / num_batches_tracked updated in place
 * (pass NULL to skip). */
size_t vg_reduce_workspace_bytes(long long rows, int channels);
size_t vg_bn_bwd_workspace_bytes(long long rows, int channels);
int vg_bn_train_fwd(const void* x, VgDType dt, long long rows, int channels, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, long long* num_batches_tracked, float momentum, float eps,
                    float* mean_out, float* rstd_out, float* scale_out, float* shift_out, float* ws, size_t ws_bytes,
                    void* stream);
/* Fused forms used by the training step: ONE cooperative launch each (grid barrier between the statistics pass and
 * the apply pass, whose re-read of the activation is served from L2).  stats = [4][channels] = mean, rstd, scale,
 * shift: written by the forward, consumed by the backward.  Shapes the fused kernel does not take fall back to the
 * two-kernel forms above internally. */
int vg_bn_act_train_fwd(const void* x, VgDType dt, long long rows, int channels, const float* gamma, const float* beta,
                        float* running_mean, float* running_var, long long* num_batches_tracked, float momentum,
                        float eps, VgAct act, float slope, float* stats, void* y, void* stream);
int vg_bn_act_train_bwd(const void* dy, const void* x, VgDType dt, long long rows, int channels, const float* stats,
                        VgAct act, float slope, float* dgamma, float* dbeta, void* dx, float* ws, size_t ws_bytes,
                        void* stream);
/* BatchNorm passes fed by the raw sums of a fused convolution epilogue (VgEpilogue above); `rows` is per group and
 * the `groups` sub-batches are consecutive in memory.  vg_bn_apply_from_sums finalises the statistics
 * (stats[groups][4][channels] written, running statistics advanced once per group, in order) and writes
 * y = act(BN(x)); vg_bn_bwd_apply_from_sums turns dz (activation derivative already applied) into the gradient of the
 * convolution output and adds dgamma / dbeta.  `rows` here are rows per group. */
int vg_bn_apply_from_sums(const void* x, VgDType dt, long long rows, int channels, int groups, const float* sums,
                          const float* gamma, const float* beta, float* running_mean, float* running_var,
                          long long* num_batches_tracked, float momentum, float eps, VgAct act, float slope,
                          float* stats, void* y, void* stream);
int vg_bn_bwd_apply_from_sums(const void* dz, const void* x, VgDType dt, long long rows, int channels, int groups,
                              const float* stats, const float* sums, float* dgamma, float* dbeta, void* dx,
                              void* stream);
/* eval-mode BatchNorm folded to scale/shift from the running statistics (main_vae.py:360, decoder.eval()). */
int vg_bn_eval_coeffs(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                      float eps, int channels, float* scale_out, float* shift_out, void* stream);
/* y = act(x * scale[c] + shift[c]); scale / shift may be NULL (identity).  Replaces the BN affine + nn.LeakyReLU /
 * nn.ReLU / nn.Sigmoid / nn.Tanh element-wise passes (main_vae.py:25,30; gan_code.py:23-50,62-85). */
int vg_scale_shift_act(const void* x, VgDType in_dt, long long rows, int channels, const float* scale,
                       const float* shift, VgAct act, float slope, void* y, VgDType out_dt, void* stream);
/* Backward of act(BN(x)):  dz = dy*act'(x*scale+shift);  dgamma += sum dz*xhat;  dbeta += sum dz;
 * dx = scale*(dz - mean(dz) - xhat*mean(dz*xhat)).  dgamma / dbeta may be NULL. */
int vg_bn_act_bwd(const void* dy, const void* x, VgDType dt, long long rows, int channels, const float* scale,
                  const float* shift, const float* mean, const float* rstd, VgAct act, float slope, float* dgamma,
                  float* dbeta, void* dx, float* ws, size_t ws_bytes, void* stream);
/* dx = dy * act'(x) for layers without BatchNorm. */
int vg_act_bwd(const void* dy, const void* x, VgDType in_dt, long long n, VgAct act, float slope, void* dx,
               VgDType out_dt, void* stream);
/* out[c] += sum_rows x[row][c]   (bias gradient of nn.Conv2d / nn.Linear, main_vae.py:23,47-48). */
int vg_colsum(const void* x, VgDType dt, long long rows, int channels, float* out, float* ws, size_t ws_bytes,
              void* stream);
/* fp32 NCHW (the reference's tensor contract, dataset_code.py:178) -> NHWC dtype.
 *   mode 0: copy;  mode 1: clamp?(src + sigma*aux)  (instance / denoising noise, vaegan_code.py:91-92,153-154);
 *   mode 2: src * (1 - aux^2)  (Tanh backward, aux = tanh output).
 * The NHWC side may carry dst_channels >= channels per pixel (extra channels are written as zeros / ignored). */
int vg_nchw_to_nhwc(const float* src, const float* aux, void* dst, VgDType dt, int batch, int channels, int h, int w,
                    int dst_channels, int mode, float sigma, int clamp, void* stream);
/* Space-to-depth image tensors (bf16 path): a <= 16-channel H x W image as [batch][H/2+o][W/2+o][64], block (Y, X),
 * sub-pixel (sy, sx), channel c at slot (sy*2+sx)*16 + c  <->  pixel (2Y-o+sy, 2X-o+sx), zeros elsewhere (o = origin =
 * the padding of the 4x4 stride-2 convolution that reads it).  In this form the image-side convolutions of the
 * reference (main_vae.py:37 first ConvBlock, gan_code.py:49 last ConvTranspose2d, gan_code.py:59 first Conv2d) are
 * ordinary 64-channel convolutions with 128-byte pixel rows.  Modes as vg_nchw_to_nhwc. */
int vg_nchw_to_s2d(const float* src, const float* aux, void* dst, int batch, int channels, int h, int w, int origin,
                   int mode, float sigma, int clamp, void* stream);
int vg_s2d_to_nchw(const void* src, float* dst, int batch, int channels, int h, int w, int origin, VgAct act,
                   float slope, void* stream);
/* uint8 NHWC image batch -> fp32 NCHW, y = (x/255 - mean)/std: the reference's ToTensor + Normalize((0.5,)*3,
 * (0.5,)*3) (dataset_code.py:147-150) on the device, so the host ships 1 byte per value instead of 4. */
int vg_u8_nhwc_to_nchw(const void* src, float* dst, int batch, int channels, int h, int w, float mean, float std,
                       void* stream);
/* nn.Linear over a flattened feature map as one dense GEMM over the NHWC activation (fc_mu / fc_logvar of
 * main_vae.py:47-48 when the feature map is large or the latent size is not a multiple of 32, e.g. the reference's own
 * 256x256 encoder: 14x14x256 -> 100).  mode 0 builds the GEMM operand W'[n_rows][kk*C] ((h, w, c) order, rows >=
 * n_valid zero) from the master W[n_valid][C*kk] ((c, h, w) order, main_vae.py:53); mode 1 adds dW' back into the
 * master-layout gradient. */
int vg_linear_permute(const float* src, float* dst, int n_valid, int n_rows, int channels, int kk, int mode,
                      void* stream);
/* dst[i] (+)= sum_{j<fan} src[idx[i*fan+j]] (negative index = no term): builds the equivalent 64-channel weights of
 * the space-to-depth convolutions from the reference-layout masters and folds their gradients back. */
int vg_gather_f32(float* dst, const float* src, const int* idx, long long n, int fan, int accumulate, void* stream);
/* NHWC dtype -> fp32 NCHW with an optional activation (Tanh of gan_code.py:50). */
int vg_nhwc_to_nchw(const void* src, VgDType dt, int src_channels, float* dst, int batch, int channels, int h, int w,
                    VgAct act, float slope, void* stream);

/* ---- VAE-GAN losses, Adam, noise ------------------------------------------------------------------------------
 * vg_reparam_fwd : logvar clamp [-10,10], std = exp(logvar/2), z = mu + std*eps, KL = -1/2 sum(1+lv-mu^2-e^lv)/B
 *                  (vaegan_code.py:75-78,114).  kl_out may be NULL.
 * vg_reparam_bwd : dmu, dlogvar from dz and the KL term weighted by kl_weight (device scalar if kl_weight_dev). */
int vg_reparam_fwd(const float* mu, const float* logvar, const float* eps, int batch, int nz, void* z, VgDType z_dt,
                   float* kl_out, void* stream);
int vg_reparam_bwd(const void* dz, VgDType dz_dt, const float* mu, const float* logvar, const float* eps, int batch,
                   int nz, const float* kl_weight_dev, float kl_weight, float* dmu, float* dlogvar, void* stream);
/* nn.BCELoss(mean) of probabilities p[n] against a constant target (vaegan_code.py:99-100,115); logs clamped at -100.
 * loss_out (=, or += when accumulate) ; dp = weight * dL/dp (may be NULL). */
int vg_bce(const float* p, int n, float target, float weight, float* loss_out, int accumulate, float* dp,
           void* stream);
/* nn.MSELoss(mean)(a, b) (vaegan_code.py:113; also the Dis_l feature-matching form of README.md:11-14 when a, b are
 * discriminator features).  grad_out = weight*2(a-b)/n (+ grad_in if given); grad pointers may be NULL. */
size_t vg_mse_workspace_bytes(void);
int vg_mse(const float* a, const float* b, long long n, float weight, const float* grad_in, float* grad_out,
           float* loss_out, void* ws, size_t ws_bytes, void* stream);
/* One launch per discriminator update (vaegan_code.py:99-101): p[2n] holds D(real) then D(fake);
 * loss_out = BCE(p[:n], target_real) + BCE(p[n:], target_fake), dp[2n] = weight * d(loss)/dp (may be NULL). */
int vg_bce_pair(const float* p, int n, float target_real, float target_fake, float weight, float* loss_out, float* dp,
                void* stream);
/* One launch for vaegan_code.py:113 + :117: loss_out = weight * MSE(a, b) (fp32 NCHW pixels, or bf16 NHWC discriminator
 * features for the Dis_l reconstruction term of README.md:11-14), grad_out = weight*2(a-b)/n (+ grad_in) in the
 * element type, and - written by the block that finishes last - total_out = loss + *w_kl_dev * *kl + w_adv * *adv
 * (NULL pointers contribute 0; total_out may be NULL).  ws: vg_mse_workspace_bytes() bytes, zeroed once by the caller;
 * the kernel re-arms it. */
int vg_mse_total(const void* a, const void* b, VgDType dt, long long n, float weight, const void* grad_in,
                 void* grad_out, float* loss_out, const float* kl, const float* adv, const float* w_kl_dev, float w_adv,
                 float* total_out, void* ws, size_t ws_bytes, void* stream);
/* total = recon + w_kl*kl + w_adv*adv  (vaegan_code.py:117); w_kl read from the device if w_kl_dev != NULL. */
int vg_total_loss(const float* recon, const float* kl, const float* adv, const float* w_kl_dev, float w_kl,
                  float w_adv, float* total, void* stream);
/* ---- data-parallel optimizer step over NVLink / NVSwitch peer memory (csrc/dp_comm.cu) -------------------------
 * The reference trains on one device (vaegan_code.py:28); its data-parallel extension averages the gradients of N
 * replicas before the three optim.Adam steps (vaegan_code.py:105,134-135).  The flat gradient / parameter buffers of
 * a network live in symmetric memory: `peer_*[r]` = rank r's buffer mapped into this process, `mc_*` = the NVSwitch
 * multicast mapping of the same buffers (both NULL: peer loads / stores instead of multimem instructions),
 * `peer_sig[r]` = rank r's signal pad (vg_dp_max_blocks() * world uint32, zeroed once), `epoch` = 2 local uint32
 * (zeroed once).  vg_dp_adam_bucket: elements [lo, lo + n) of the buffers - reduce-scatter of the gradients, Adam on
 * this rank's 1/world slice (m, v: this rank's moment buffers, full length, only the slice is touched), all-gather
 * of the new parameters, in ONE kernel.  Every rank must launch the same sequence of calls.  *step_dev as in
 * vg_adam_step but NOT incremented here (the caller ticks it once per optimizer step).  write_grads != 0 also
 * leaves the summed gradient in every replica (parity tests).  *err_flag is set if a cross-GPU barrier times out. */
#define VG_DP_MAX_RANKS 8
typedef struct VgDpComm {
    float* mc_grads;
    float* mc_params;
    float* peer_grads[VG_DP_MAX_RANKS];
    float* peer_params[VG_DP_MAX_RANKS];
    unsigned int* peer_sig[VG_DP_MAX_RANKS];
    unsigned int* epoch;
    int rank;
    int world;
} VgDpComm;
int vg_dp_max_blocks(void);
int vg_dp_adam_bucket(const VgDpComm* comm, long long lo, long long n, float* m, float* v, double lr, double beta1,
                      double beta2, double eps, const long long* step_dev, float grad_scale, int blocks,
                      int write_grads, int* err_flag, void* stream);
/* *step_dev += 1 (the tick vg_adam_step does itself). */
int vg_adam_tick(long long* step_dev, void* stream);
/* vg_adam_step without the tick: the update of one RANGE of a flat buffer (pointers already offset), for optimizer
 * steps issued bucket by bucket from inside the backward pass; the caller ticks *step_dev once per step first. */
int vg_adam_apply(float* p, const float* g, float* m, float* v, long long n, double lr, double beta1, double beta2,
                  double eps, const long long* step_dev, float grad_scale, void* stream);
/* torch.optim.Adam.step (vaegan_code.py:42-44,105,134-135) over one flat fp32 buffer; *step_dev is incremented
 * first and drives the bias corrections, so the call can be replayed from a CUDA graph.  g is multiplied by
 * grad_scale (1/world_size under data parallelism). */
int vg_adam_step(float* p, const float* g, float* m, float* v, long long n, double lr, double beta1, double beta2,
                 double eps, long long* step_dev, float grad_scale, void* stream);
/* Standard normal noise (torch.randn_like, vaegan_code.py:77,91,92): Philox4x32-10 + Box-Muller, keyed by
 * (seed, *offset_dev, stream_id); *offset_dev is incremented after the launch when given. */
int vg_randn(float* out, long long n, unsigned long long seed, unsigned long long* offset_dev,
             unsigned long long stream_id, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VAEGAN_B200_H */
