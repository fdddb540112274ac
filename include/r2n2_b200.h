/*
 * r2n2_b200.h -- C ABI of the B200-native 3D-R2N2 hot path (libr2n2_b200.so).
 *
 * The reference (chrischoy/3D-R2N2) has no FFI: its layers call Theano ops inline.  Each
 * entry point below therefore replaces one Theano call site of lib/layers.py /
 * lib/solver.py / lib/voxel.py (cited per function, paths relative to the reference root);
 * the Python mirror of the reference's layer API (3d-r2n2_b200/lib/layers.py) binds them
 * through ctypes exactly as INTEGRATION.md shows.
 *
 * Conventions
 *   - plain pointers + sizes only; every pointer is a DEVICE pointer unless stated.
 *   - the library never allocates, frees or retains caller memory; all work is enqueued on
 *     the caller's stream (cudaStream_t passed as void*); no implicit synchronisation.
 *   - every function returns 0 on success or a negative R2N2_ERR_* code; the message is
 *     available from r2n2_last_error() (thread-local).
 *   - activations are CHANNELS-LAST: 2-D tensors [N,H,W,C], 3-D tensors [B,D,H,W,C]
 *     (a 2-D tensor is the D==1 case).  The Python layer exposes them with the reference's
 *     logical shapes ((N,C,H,W) / (B,D,C,H,W)) as stride-permuted views.
 *   - dtype codes: R2N2_F32 = 0, R2N2_BF16 = 1.
 *   - convolution weights are PACKED: w[Cout][kd][kh][kw][Cin], already flipped so that the
 *     kernel computes a correlation  y[p,co] = sum_{tap,ci} x[p+tap-centre,ci] * w[co,tap,ci]
 *     (Theano's conv2d / conv3d2d.conv3d are true convolutions; the flip happens once, at
 *     Weight.set_value time, see 3d-r2n2_b200/lib/layers.py:pack_*).
 */
#ifndef R2N2_B200_H
#define R2N2_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define R2N2_F32 0
#define R2N2_BF16 1

#define R2N2_OK 0
#define R2N2_ERR_BAD_SHAPE (-1)
#define R2N2_ERR_BAD_DTYPE (-2)
#define R2N2_ERR_BAD_ALIGN (-3)
#define R2N2_ERR_WORKSPACE (-4)
#define R2N2_ERR_CUDA (-5)
#define R2N2_ERR_UNSUPPORTED (-6)

/* conv epilogue flags: v = acc; BIAS: v += bias[co]; LRELU: v = v>0?v:0.01v;
 * RESIDUAL: v += residual[p,co]; MASK: v *= (mask_src[p,co] > 0 ? 1 : 0.01)   (in that order) */
#define R2N2_EPI_BIAS 1
#define R2N2_EPI_LRELU 2
#define R2N2_EPI_RESIDUAL 4
#define R2N2_EPI_MASK 8
#define R2N2_EPI_RES_PRE 64  /* with R2N2_EPI_RESIDUAL: the residual is added to the accumulator BEFORE bias / LeakyReLU
                                (a partial sum of the same contraction, e.g. the hi/lo operand-split passes of
                                COMPUTE='bf16x3'), instead of after the activation */
/* geometry flags of r2n2_conv_fwd (R2N2_ALGO_UMMA only), N,D,H,W always name the SMALL grid:
 * UP2  : fused Unpool3DLayer + Conv3DLayer (lib/layers.py:375-385 + 441): x is [N,D,H,W,Cin], the
 *        convolution runs over its zero-stuffed x2 upsampling and y / residual / mask_src are
 *        [N,2D,2H,2W,Cout]; only the taps that meet a stored value are computed (1/8 of the MACs).
 * DOWN2: x is [N,2D,2H,2W,Cin], y[N,D,H,W,Cout] holds the even positions of the 'same' convolution
 *        (with transposed weights: the gradient of an UP2 convolution w.r.t. its small input). */
#define R2N2_CONV_UP2 16
#define R2N2_CONV_DOWN2 32
/* r2n2_conv_wgrad `accumulate` argument is a bit set: ACCUMULATE adds into dw; UP2: x is the small
 * [N,D,H,W,Cin] tensor of an UP2 convolution and dy is [N,2D,2H,2W,Cout]. */
#define R2N2_WGRAD_ACCUMULATE 1
#define R2N2_WGRAD_UP2 2

/* conv algorithms */
#define R2N2_ALGO_SIMT 0  /* fp32 FFMA implicit GEMM: the exact-fp32 parity path            */
#define R2N2_ALGO_UMMA 1  /* tcgen05 + TMA implicit GEMM, bf16 operands, fp32 TMEM accumulate */

const char* r2n2_version(void);
const char* r2n2_last_error(void);

/* ---- dense contractions -------------------------------------------------------------
 * r2n2_conv_fwd: 'same' stride-1 N-d convolution as an implicit GEMM (M = N*D*H*W pixels,
 * N = Cout, K = kd*kh*kw*Cin).  Replaces tensor.nnet.conv2d (lib/layers.py:322-333),
 * conv3d2d.conv3d (lib/layers.py:441-442, 504) and, with D=H=W=1, k=1, tensor.dot
 * (lib/layers.py:159-161, 502-503).  The same entry point computes data gradients when
 * given the transposed/flipped weights from r2n2_weight_transpose (tensor.grad,
 * models/net.py:58).  mask_src / residual / y have dtype out_dtype; x / w have in_dtype. */
int r2n2_conv_fwd(const void* x, const void* w, const float* bias, const void* mask_src,
                  const void* residual, void* y, int N, int D, int H, int W, int Cin, int Cout,
                  int kd, int kh, int kw, int flags, int in_dtype, int out_dtype, int algo,
                  void* stream);

/* r2n2_conv_fwd_colsum: r2n2_conv_fwd plus colsum[c] += sum over output pixels of y[p,c] (fp32, always
 * accumulates).  Used for data gradients: y is then the gradient of the producing layer's output and
 * colsum its bias gradient (`+ b.dimshuffle(...)`, lib/layers.py:333,442) -- fused into the TMA epilogue
 * where the tile is in shared memory anyway, otherwise one r2n2_bias_grad pass over y. */
int r2n2_conv_fwd_colsum(const void* x, const void* w, const float* bias, const void* mask_src,
                         const void* residual, void* y, int N, int D, int H, int W, int Cin, int Cout,
                         int kd, int kh, int kw, int flags, int in_dtype, int out_dtype, int algo,
                         float* colsum, void* stream);

/* r2n2_conv_fwd_gru (tcgen05 path, bf16 tensors): a hidden-state convolution of the conv-GRU (FCConv3DLayer's
 * `conv3d(h, Wh)`, lib/layers.py:504) with the gate math of models/res_gru_net.py:130-151 -- or of its gradient
 * (tensor.grad, models/net.py:58) -- applied to the accumulator in the epilogue instead of a separate gate
 * kernel; no convolution output is written.  pr = hidden-grid position (row of [N*D*H*W]), HC = hidden channels:
 *   R2N2_GRU_UR    x = h_prev, Cout = 2 HC: s = sigmoid(acc + in0[pr,co] + bias[co])   (in0 = x-projection, [.,2HC])
 *                  co < HC: out0[pr,co] = s (update gate); else c = co-HC: out1[pr,c] = s (reset gate),
 *                  out2[pr,c] = s * in1[pr,c]   (in1 = h_prev)
 *   R2N2_GRU_BLEND x = r*h, Cout = HC: t = tanh(acc + in0[pr,c] + bias[c]); out0 = t;
 *                  out1 = in1 * in2 + (1 - in1) * t   (in1 = update gate, in2 = h_prev; out1 = h_new)
 *   R2N2_GRU_RBWD  x = d(pre-tanh), w = Wh_c transposed, Cout = HC: g = acc; out0[pr, HC + c] = g * in0 * in1 * (1 - in1)
 *                  (in0 = h_prev, in1 = reset gate; out0 = d(pre-sigmoid) [.,2HC]); out1[pr,c] += g * in1   (dh_prev)
 *   R2N2_GRU_OBWD  x = d(pre-sigmoid) [.,2HC], w = Wh_ur transposed, Cout = HC: dh = acc + in0[pr,c]; with u = in1,
 *                  cand = in2, hp = in3 (or 0 when NULL) of the EARLIER step: out0 = dh (1-u)(1-cand^2),
 *                  out1[pr,c] = dh (hp - cand) u (1-u)  ([.,2HC] rows), out2 = dh * u.
 * The standalone gate kernels (r2n2_gru_gates_*) remain for the first view (h = 0: no convolution) and the fp32 path. */
#define R2N2_GRU_UR 1
#define R2N2_GRU_BLEND 2
#define R2N2_GRU_RBWD 3
#define R2N2_GRU_OBWD 4
int r2n2_conv_fwd_gru(int mode, const void* x, const void* w, const float* bias, const void* in0, const void* in1,
                      const void* in2, const void* in3, void* out0, void* out1, void* out2, int N, int D, int H,
                      int W, int Cin, int Cout, int kd, int kh, int kw, int dtype, void* stream);

/* r2n2_conv_wgrad: dw[co][tap][ci] (+)= sum_p dy[p,co] * x[p+tap-centre,ci]  (fp32 output in
 * packed layout).  tensor.grad of the call sites above w.r.t. W / Wh / Wx. */
int r2n2_conv_wgrad(const void* x, const void* dy, float* dw, int N, int D, int H, int W, int Cin,
                    int Cout, int kd, int kh, int kw, int dtype, int algo, int accumulate,
                    void* stream);

/* r2n2_bias_grad: db[c] (+)= sum_rows dy[row,c].  Gradient of the `+ b.dimshuffle(...)`
 * terms (lib/layers.py:161,333,442,505). */
int r2n2_bias_grad(const void* dy, float* db, long long rows, int C, int dtype, int accumulate,
                   void* stream);

/* w[Cout][taps][Cin] (fp32 master) -> wt[Cin][taps reversed][Cout] in out_dtype: the operand
 * r2n2_conv_fwd needs to compute the data gradient. */
int r2n2_weight_transpose(const float* w, void* wt, int Cout, int taps, int Cin, int out_dtype,
                          void* stream);

/* the same for every weight tensor of a network in one launch (after each optimiser step): `table`
 * is a DEVICE array of n_entries rows of 6 int64 {src offset in floats from base, dst offset in
 * elements from wt_base (even), Cout, taps, Cin, index of the entry's first tile}; a tile is 64 output
 * channels x 32 input channels of one tap: total_tiles = sum over entries of ceil(Cin/32) * ceil(Cout/64) * taps. */
int r2n2_weight_transpose_batch(const float* base, void* wt_base, const long long* table, int n_entries,
                                long long total_tiles, int out_dtype, void* stream);

/* The same from the bf16 operand mirror (same layout / element offsets as the fp32 master buffer; every segment
 * offset, Cin and Cout even): table[e][5] = first tile of the entry under a 64 (co) x 64 (ci) tiling, i.e.
 * ceil(Cin/64) * ceil(Cout/64) * taps tiles per entry.  Bit-identical output, half the read traffic. */
int r2n2_weight_transpose_batch_bf16(const void* mirror_base, void* wt_base, const long long* table, int n_entries,
                                     long long total_tiles, void* stream);

/* fp32 -> out_dtype copy (bf16 operand mirrors of the fp32 master weights). */
int r2n2_cast(const float* src, void* dst, long long n, int out_dtype, void* stream);

/* fp32 -> (hi, lo) bf16 pair with src = hi + lo up to 2^-17 relative: the operand split of the
 * COMPUTE='bf16x3' path (three tcgen05 passes hi*hi + lo*hi + hi*lo accumulated in fp32 reach the
 * reference's float32 accuracy, .theanorc:2, at a third of the bf16 tensor rate). */
int r2n2_split_bf16(const float* src, void* hi, void* lo, long long n, void* stream);

/* ---- bandwidth-bound kernels ----------------------------------------------------------
 * input images: x (N,C,H,W) fp32 as delivered by lib/data_process.py:133-148 -> channels-last
 * [N,H,Wpad,Cpad] in out_dtype; pixels are written at column offset xoff, everything else is
 * zero (Wpad=W, xoff=0 for the plain case). */
int r2n2_nchw_to_nhwc(const float* x, void* y, int N, int C, int H, int W, int Cpad, int Wpad,
                      int xoff, int out_dtype, void* stream);

/* first-layer variant: y[n,h,w, kx*C + c] = x[n,c,h, w + kx - (kw-1)/2], zero outside the image and
 * for channels > kw*C, so that the kh x kw conv1 (models/res_gru_net.py:33) becomes a kh x 1
 * convolution over Cpack channels (weights packed accordingly by the host mirror).  If kw*C < Cpack, channel
 * kw*C of every pixel holds 1.0 (round 2): with zero weights there the forward pass is unchanged, and the weight
 * gradient of the centre tap at that channel is sum_p dy[p, co] -- the layer's BIAS gradient, for free. */
int r2n2_nchw_to_rowpack(const float* x, void* y, int N, int C, int H, int W, int kw, int Cpack,
                         int out_dtype, void* stream);

/* PoolLayer (lib/layers.py:349-353): 2x2/2 max pool, pad 1 (-inf), floor.  If b != NULL the
 * pooled tensor is a+b (the AddLayer that precedes the pool in models/res_gru_net.py:92-93). */
int r2n2_maxpool2d_fwd(const void* a, const void* b, void* y, int N, int H, int W, int C, int dtype,
                       void* stream);
/* gradient to every maximal element of each (non-overlapping) window, as Theano does;
 * lrelu_mask: additionally multiply by LeakyReLU'(a) (a is then a LeakyReLU output). */
int r2n2_maxpool2d_bwd(const void* a, const void* b, const void* dy, void* dx, int N, int H, int W,
                       int C, int lrelu_mask, int dtype, void* stream);

/* The same backward pass with what follows it in the graph fused in (each pointer optional):
 * dxb      = dx * LeakyReLU'(b): the gradient w.r.t. the pre-activation of the branch b = LeakyReLU(z_b) of
 *            pool(a + b) (res_gru_net.py:100-140: a = 1x1 shortcut, b = rectified 3x3 branch); needs b;
 * colsum_a = fp32 [C] += column sums of dx (after lrelu_mask), colsum_b += column sums of dxb: the bias
 *            gradients of the convolutions that produced a and b (atomic adds: zero or accumulate-into). */
int r2n2_maxpool2d_bwd_fused(const void* a, const void* b, const void* dy, void* dx, void* dxb, float* colsum_a,
                             float* colsum_b, int N, int H, int W, int C, int lrelu_mask, int dtype, void* stream);

/* Training variant of the pair above that does not re-read the pooled tensors in the backward pass:
 * r2n2_maxpool2d_fwd_code additionally writes one byte per OUTPUT element, code[N,Ho,Wo,C]: bit k (k = 2 dh + dw for
 * the window position (2 ho - 1 + dh, 2 wo - 1 + dw)) = position k attains the maximum (every tied position, as
 * above); bit 4 + k = the mask source -- b if given, else a -- is > 0 at position k.
 * r2n2_maxpool2d_bwd_code computes the same dx / dxb / colsum_a / colsum_b as r2n2_maxpool2d_bwd_fused from dy and the
 * code alone (bit-identical results): lrelu_mask uses the recorded sign and is therefore only valid for a code
 * recorded WITHOUT b, dxb only for a code recorded WITH b (both at once: R2N2_ERR_UNSUPPORTED).  HBM traffic of the
 * backward pass: dy + code + dx (+ dxb) instead of a (+ b) + dy + dx (+ dxb). */
int r2n2_maxpool2d_fwd_code(const void* a, const void* b, void* y, unsigned char* code, int N, int H, int W, int C,
                            int dtype, void* stream);
int r2n2_maxpool2d_bwd_code(const unsigned char* code, const void* dy, void* dx, void* dxb, float* colsum_a,
                            float* colsum_b, int N, int H, int W, int C, int lrelu_mask, int dtype, void* stream);

/* Unpool3DLayer (lib/layers.py:375-385): y[B,2D,2H,2W,C], value at the even corner, zeros
 * elsewhere; input is a+b if b != NULL (AddLayer res7/res8, models/res_gru_net.py:171-178). */
int r2n2_unpool3d_fwd(const void* a, const void* b, void* y, int B, int D, int H, int W, int C,
                      int dtype, void* stream);
int r2n2_unpool3d_bwd(const void* dy, void* dx, int B, int D, int H, int W, int C, int dtype,
                      void* stream);

/* AddLayer (lib/layers.py:214) and LeakyReLU backward (lib/layers.py:641-643):
 * y = a + b ;  dx = dy * (a > 0 ? 1 : 0.01) (+ addend if not NULL). */
int r2n2_add(const void* a, const void* b, void* y, long long n, int dtype, void* stream);
int r2n2_lrelu_fwd(const void* x, void* y, long long n, int dtype, void* stream);
int r2n2_lrelu_bwd(const void* a, const void* dy, const void* addend, void* dx, long long n,
                   int dtype, void* stream);
/* The same over a [rows, C] tensor, plus colsum[c] += sum over rows of dx[row, c] (fp32, the values as stored): dx is
 * the gradient of a convolution's pre-activation, so colsum is that convolution's bias gradient (`+ b.dimshuffle(...)`,
 * lib/layers.py:442) -- no separate r2n2_bias_grad pass over dx. */
int r2n2_lrelu_bwd_colsum(const void* a, const void* dy, const void* addend, void* dx, long long rows, int C,
                          float* colsum, int dtype, void* stream);

/* SigmoidLayer / TanhLayer / ComplementLayer (1-x) / EltwiseMultiplyLayer (a*b) as standalone
 * kernels (lib/layers.py:649-676, 217-225) for the layer-by-layer API; the fused GRU kernels
 * below are what the networks use. */
#define R2N2_ELT_SIGMOID 0
#define R2N2_ELT_TANH 1
#define R2N2_ELT_COMPLEMENT 2
#define R2N2_ELT_MUL 3
/* their derivatives (tensor.grad through a user-defined network_definition, models/net.py:56-58):
 * SIGMOID_BWD: y = b * a * (1 - a) with a = the sigmoid OUTPUT, b = dL/d(output); TANH_BWD: y = b * (1 - a^2);
 * NEG: y = -a (ComplementLayer). */
#define R2N2_ELT_SIGMOID_BWD 4
#define R2N2_ELT_TANH_BWD 5
#define R2N2_ELT_NEG 6
int r2n2_eltwise(int op, const void* a, const void* b, void* y, long long n, int dtype,
                 void* stream);

/* 3-D conv GRU gate math (models/res_gru_net.py:130-151; lib/layers.py:217-225,649-676).
 * Hidden grid h: [B,64,128] (= (B,4,4,4,128) channels-last).
 *   ur_fwd : u = sigmoid(conv_ur[...,0:128] + xp_ur[...,0:128] + b[0:128]),
 *            r = sigmoid(conv_ur[...,128:256] + xp_ur[...,128:256] + b[128:256]), rh = r*h
 *            (conv_ur may be NULL: first step, h == 0)
 *   out_fwd: c = tanh(conv_c + xp_c + b_c);  h_new = u*h + (1-u)*c
 *   out_bwd: da_c = dh*(1-u)*(1-c^2); da_u = dh*(h-c)*u*(1-u) -> da_ur[...,0:128];
 *            dh_prev = dh*u
 *   r_bwd  : da_r = d_rh*h*r*(1-r) -> da_ur[...,128:256];  dh_prev += d_rh*r            */
int r2n2_gru_gates_ur_fwd(const void* conv_ur, const void* xp_ur, const float* bias_ur,
                          const void* h, void* u, void* r, void* rh, int B, int dtype,
                          void* stream);
int r2n2_gru_gates_out_fwd(const void* conv_c, const void* xp_c, const float* bias_c, const void* u,
                           const void* h, void* c, void* h_new, int B, int dtype, void* stream);
int r2n2_gru_gates_out_bwd(const void* dh, const void* u, const void* c, const void* h, void* da_ur,
                           void* da_c, void* dh_prev, int B, int dtype, void* stream);
int r2n2_gru_gates_r_bwd(const void* d_rh, const void* h, const void* r, void* da_ur,
                         void* dh_prev, int B, int dtype, void* stream);

/* 3-D conv LSTM variant (no reference code: lib/layers.py:508-575 is a stub; built from
 * imgs/lstm.png / README.md:36-38): gates a = conv + xp + b laid out [B,64,384] = (f,i,g):
 *   f = sigmoid(a_f), i = sigmoid(a_i), g = tanh(a_g); s_new = f*s + i*g; h_new = tanh(s_new)
 *   bwd: ds_tot = dh*(1-h_new^2) + ds_in; da_f = ds_tot*s*f(1-f); da_i = ds_tot*g*i(1-i);
 *        da_g = ds_tot*i*(1-g^2); ds_prev = ds_tot*f                                      */
int r2n2_lstm_gates_fwd(const void* conv_fig, const void* xp_fig, const float* bias_fig,
                        const void* s, void* fig, void* s_new, void* h_new, int B, int dtype,
                        void* stream);
int r2n2_lstm_gates_bwd(const void* dh, const void* ds_in, const void* fig, const void* s,
                        const void* h_new, void* da_fig, void* ds_prev, int B, int dtype,
                        void* stream);

/* SoftmaxWithLoss3D (lib/layers.py:583-601), fused forward(+backward):
 *   logits  [B,nv,nv,nv,2] fp32 channels-last (conv11 output)
 *   y       (B,nv,2,nv,nv) fp32 reference layout, may be NULL (lib/solver.py:222-226 feeds 0)
 *   pred    (B,nv,2,nv,nv) fp32 reference layout, may be NULL  = exp(x)/sum_c exp(x)
 *   loss_sum: 1 float, ACCUMULATES sum_vox(-sum_c y_c x_c + log sum_c exp x_c), may be NULL
 *   dlogits [B,nv,nv,nv,2] in ddtype, may be NULL                = (p - y) * grad_scale
 *   cstride : channels per voxel in logits / dlogits (2, or 16 when conv11 is padded to a
 *             tensor-core N granule; the extra channels are ignored / written as zero)      */
int r2n2_softmax_ce3d(const float* logits, const float* y, float* pred, float* loss_sum,
                      void* dlogits, int B, int nv, float grad_scale, int ddtype, int cstride,
                      void* stream);

/* ADAM (lib/solver.py:20-48) over one flat fp32 buffer; the first n_decay elements are
 * non-bias parameters (L2 term g += wd*p), the rest biases.  g is multiplied by grad_scale
 * first (1/world after the NCCL sum).  p_mirror (may be NULL) receives p in mirror_dtype. */
int r2n2_adam(float* p, const float* g, float* m, float* v, void* p_mirror, long long n,
              long long n_decay, float lr_t, float beta1, float beta2, float eps, float wd,
              float grad_scale, int mirror_dtype, void* stream);
/* SGD with momentum (lib/solver.py:51-71) */
int r2n2_sgd(float* p, const float* g, float* vel, void* p_mirror, long long n, long long n_decay,
             float lr, float momentum, float wd, float grad_scale, int mirror_dtype, void* stream);

/* evaluate_voxel_prediction (lib/voxel.py:4-11): pred, gt (B,nv,2,nv,nv) fp32; counts[B][5] =
 * {xor, intersection, union, false-pos, false-neg} at pred[:,:,1] >= thresh (int64, zeroed here) */
int r2n2_voxel_eval(const float* pred, const float* gt, float thresh, long long* counts, int B,
                    int nv, void* stream);

/* Input pipeline (lib/data_augmentation.py:6-70, preprocess_img): rgba [n_img,Hs,Ws,4] uint8 renders;
 * params [n_img][6] int32 = {crop row, crop column, flip, bg r, bg g, bg b} (the caller draws them in the
 * reference's order; the kernel does not check that the crop window lies inside the source).  Pixels with
 * alpha == 0 take the background colour, the (Ho,Wo) window is cropped, optionally mirrored, divided by
 * 255 in fp32 and written channels-first: out [n_img,3,Ho,Wo] in out_dtype (= batch_img of
 * lib/data_process.py:145-148, viewed as (T,B,3,H,W) by the caller). */
int r2n2_preprocess_views(const unsigned char* rgba, const int* params, void* out, int n_img, int Hs, int Ws,
                          int Ho, int Wo, int out_dtype, void* stream);

/* Voxel labels of one batch (lib/data_process.py:135-154): occ [B,nv,nv,nv] bytes, 0 = empty, 1 = occupied
 * (the boolean grid read from model.binvox), any other value = unfilled slot of a short last batch;
 * y [B,nv,2,nv,nv] fp32 with y[:,:,0] = (occ == 0), y[:,:,1] = (occ == 1): unfilled slots are zero in both
 * channels, as the reference leaves them. */
int r2n2_voxel_labels(const unsigned char* occ, float* y, int B, int nv, void* stream);

/* Per-sample average precision of the occupancy channel (lib/test_net.py:69-70:
 * sklearn.metrics.average_precision_score(gt[j,:,1].flatten(), pred[j,:,1].flatten())): pred, gt
 * (B,nv,2,nv,nv) fp32, nv^3 <= 32768; ap[B] float64 = sum over distinct score values (descending) of
 * (recall - previous recall) * precision, 0 when a sample has no positive voxel. */
int r2n2_average_precision(const float* pred, const float* gt, double* ap, int B, int nv, void* stream);

int r2n2_fill_zero(void* p, long long bytes, void* stream);

/* 1 if any tcgen05 kernel launched by this process hit its bounded mbarrier wait (a pipeline
 * protocol failure: results of that launch are garbage); synchronises the device. */
int r2n2_umma_error_flag(void);

/* Upper bound on the SMs the persistent (one CTA per SM) grids of later launches occupy; 0 = all.  The data-parallel
 * trainer lowers it while a gradient all-reduce is in flight: a 148-CTA grid that finds SMs taken by the NCCL kernels
 * runs its last CTAs as a second wave.  Process-wide, takes effect at the next launch.  Returns R2N2_OK. */
int r2n2_set_sm_limit(int n_sms);

#ifdef __cplusplus
}
#endif
#endif /* R2N2_B200_H */
