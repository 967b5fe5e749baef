/* spectre_mix.h -- C ABI of the B200-native Spectre spectral-mix forward path.
 *
 * One shared library (fft_b200/_C/libspectre_mix.so), plain pointers and sizes,
 * no torch types.  Every entry point states which lines of the reference
 * (/root/reference/spectre.py) it stands in for; the reference itself has no
 * FFI layer -- it calls torch.fft from Python -- so these are the symbols a
 * binding (ctypes / pybind / torch.library) attaches to.  See INTEGRATION.md.
 *
 * Function computed by the forward entry points, for b < B, n < min(N, n_fft), c < C:
 *
 *   out[b, n, c] = irfft_{n_fft}( gate[b, c / group_width, :] * rfft_{n_fft}(V[b, :, c]) + mem[:, c] )[n]
 *
 * with rfft zero-padding (N < n_fft) or truncating (N > n_fft) its input, irfft
 * scaled by 1/n_fft and ignoring the imaginary parts of bin 0 and bin n_fft/2.
 * All heads of a SpectreMultiHead go in ONE call: C = embed_dim, the gate rows of
 * head h are rows [h*G, (h+1)*G) of `gate`, group_width = head_dim / G.
 *
 * Threading / ownership: the caller owns every buffer; calls are asynchronous on
 * `stream` (device entry points) and never synchronise the device.  The only
 * library-owned device memory is (a) one twiddle table per (device, n_fft),
 * built and uploaded on the FIRST call with that n_fft (a synchronous copy: make
 * one warm-up call before capturing a CUDA graph), read lock-free afterwards,
 * and (b) for callers of spectre_mix_fwd that pass no workspace at
 * n_fft > 4096, a stream-ordered allocation (cudaMallocFromPoolAsync /
 * cudaFreeAsync on the caller's stream, private pool) per call -- never shared
 * between streams, legal under stream capture.  Pass a workspace
 * (spectre_mix_workspace_bytes + spectre_mix_fwd_ws) to own that memory too.
 * Launches take no lock: host threads driving different streams do not
 * serialise on the library.  There is no CPU fallback and no cuFFT: an
 * unsupported argument returns an error code and sets spectre_mix_last_error().
 */
#ifndef SPECTRE_MIX_H_
#define SPECTRE_MIX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element types of V and out */
#define SPECTRE_MIX_F32 0
#define SPECTRE_MIX_BF16 1

/* return codes (0 = ok; CUDA runtime errors are returned as 1000 + cudaError_t) */
#define SPECTRE_MIX_OK 0
#define SPECTRE_MIX_ERR_BAD_ARG 1        /* null pointer, negative size, C % group_width != 0 ... */
#define SPECTRE_MIX_ERR_UNSUPPORTED 2    /* n_fft not a power of two in [32, 16384], misaligned pointer/stride */
#define SPECTRE_MIX_ERR_NO_DEVICE 3
#define SPECTRE_MIX_ERR_CUDA 1000

/* ABI version of this header; bumped on any signature change. */
int spectre_mix_abi_version(void);

/* Thread-local description of the last non-zero return code on this host thread. */
const char *spectre_mix_last_error(void);

/* Fused rfft -> gate multiply (+ memory) -> irfft -> [:N] on device buffers.
 * Replaces spectre.py:506 (rfft), :542-545 (gate broadcast + complex multiply),
 * :548-549 (memory add), :551 (irfft), :553 (slice) and the per-head loop /
 * torch.cat of :703-718.
 *
 *   v          device, [B][N][C], element stride 1 along C; strides in elements
 *   gate       device, complex64 [B][C/group_width][n_fft/2+1] contiguous
 *   mem        device or NULL, complex64 [n_fft/2+1][C], row stride mem_stride (complex elements)
 *   out        device, [B][min(N,n_fft)][C], element stride 1 along C
 *   stream     cudaStream_t (NULL = legacy default stream)
 */
int spectre_mix_fwd(const void *v, int v_dtype, int64_t v_stride_b, int64_t v_stride_n,
                    const void *gate, const void *mem, int64_t mem_stride,
                    void *out, int out_dtype, int64_t out_stride_b, int64_t out_stride_n,
                    int B, int N, int n_fft, int C, int group_width, void *stream);

/* Scratch the call needs for the given problem (bytes; 0 for every n_fft <= 4096 and for layouts that run as a
 * single kernel).  n_fft = 8192 / 16384 with a packed layout run as three launches around a complex intermediate
 * [B][n_fft][C] fp32 that lives in the workspace (the kernel never allocates). */
size_t spectre_mix_workspace_bytes(int v_dtype, int B, int N, int n_fft, int C, int group_width);

/* spectre_mix_fwd with caller-owned scratch: `workspace` = device memory of at least spectre_mix_workspace_bytes(...)
 * bytes, 16-byte aligned, used only by this call's launches on `stream` (so one workspace per stream in flight).
 * workspace = NULL behaves like spectre_mix_fwd (stream-ordered internal allocation when scratch is needed). */
int spectre_mix_fwd_ws(const void *v, int v_dtype, int64_t v_stride_b, int64_t v_stride_n,
                       const void *gate, const void *mem, int64_t mem_stride,
                       void *out, int out_dtype, int64_t out_stride_b, int64_t out_stride_n,
                       int B, int N, int n_fft, int C, int group_width,
                       void *workspace, size_t workspace_bytes, void *stream);

/* The same mix with the gate generator's tail FUSED into the kernel: instead of a materialised gate the call
 * takes the anchors and evaluates, inside the kernel's gate staging, for the gate row g of batch row b at bin k
 *   gate[b, g, k] = modReLU_g( cubic_interp(planes of anchors[b, head(g)])(k) ) * pos_phase[b or 0, k]
 * (spectre.py:526-528 -> :26-61 grid_sample bicubic / border / align_corners, :531 -> :109-121, :534-536; arguments as
 * spectre_gate_expand below), then proceeds as spectre_mix_fwd (spectre.py:506, :542-553).  The (B, NG, F_half) gate tensor
 * is never written or read: algorithmic bytes drop to V + out + anchors + bias.  Layouts the packed kernels cannot take
 * (group_width not a multiple of 4, misaligned rows) expand the gate into the workspace first and run the plain kernels.
 * workspace >= spectre_mix_anchors_workspace_bytes(...) or NULL (stream-ordered internal allocation when needed). */
size_t spectre_mix_anchors_workspace_bytes(int v_dtype, int B, int N, int n_fft, int C, int group_width);
int spectre_mix_fwd_anchors(const void *v, int v_dtype, int64_t v_stride_b, int64_t v_stride_n,
                            const void *anchors, const float *bias, const float *eps,
                            const void *pos_phase, int64_t pos_stride_b, int G, int Bk,
                            const void *mem, int64_t mem_stride,
                            void *out, int out_dtype, int64_t out_stride_b, int64_t out_stride_n,
                            int B, int N, int n_fft, int C, int group_width,
                            void *workspace, size_t workspace_bytes, void *stream);

/* Gradient of the mix with respect to the gate (autograd through spectre.py:506-553), for dY = d loss / d out:
 *   dgate[b, g, k] = w_k / n_fft * sum over the group's channels c of conj(rfft(V[b,:,c]))[k] * rfft(dY[b,:,c])[k]
 * (w = 1 at bin 0 and n_fft/2, 2 elsewhere; imaginary parts at those two bins are zero, as irfft ignores them).  ONE kernel:
 * V's tile is transformed and its packed spectrum parked in tensor memory, dY's tile of the same channels follows, the
 * products are reduced over channel pairs, the mirrored bin and the gate group on chip -- neither spectrum reaches HBM.
 *   v, dy   device, [B][N][C] (dy: [B][min(N,n_fft)][C] is fine, rows beyond are zero), same dtype, strides in elements
 *   dgate   device, complex64 [B][C/group_width][n_fft/2+1] contiguous (overwritten)
 * Built for n_fft = 4096, group widths that are multiples of 8 and 16-byte aligned rows; returns SPECTRE_MIX_ERR_UNSUPPORTED
 * otherwise (the binding then forms the two spectra with spectre_rfft_fwd).  The gradient with respect to V needs no entry
 * of its own: dV = spectre_mix_fwd(dY, conj(gate)). */
int spectre_mix_dgate(const void *v, const void *dy, int dtype, int64_t v_stride_b, int64_t v_stride_n,
                      int64_t dy_stride_b, int64_t dy_stride_n, void *dgate,
                      int B, int N, int n_fft, int C, int group_width, void *stream);

/* Same function on HOST buffers (what a non-CUDA host language binds): copies
 * v/gate/mem to the device in batch chunks, runs the kernel and copies the
 * result back, overlapping the three on internal streams; returns when `out`
 * is complete.  Contiguous layouts, fp32 only.  Pinned host memory is used
 * as-is; pageable memory works but copies slower. */
int spectre_mix_fwd_host(const float *v, const float *gate, const float *mem, float *out,
                         int B, int N, int n_fft, int C, int group_width);

/* The host entry keeps its device staging (four streams, each with a V, an out and a gate chunk of up to ~100 MB of V:
 * ~0.85 GB at seq 4096 x d 768, plus the long-context workspaces) between calls so that a steady caller never allocates.
 * This frees it for the current device; the next spectre_mix_fwd_host call allocates again.  Returns 0. */
int spectre_mix_host_release(void);

/* Introspection (no GPU needed): the batch-chunk schedule spectre_mix_fwd_host uses for B rows of N x C fp32 -- rows per chunk,
 * ramping up by doubling from ~12 MB of V to ~100 MB and back down (short pipeline fill and drain, few hand-overs in between).
 * Writes at most `cap` entries to rows_out and returns the number of chunks (-1 on bad arguments). */
int spectre_mix_host_schedule(int B, int N, int C, int *rows_out, int cap);

/* Forward half only: spec[b, k, c] = rfft_{n_fft}(V[b, :, c])[k], k <= n_fft/2.
 * Replaces spectre.py:776-777 (PrefixFFTCache.prefill: pad + rfft along dim 0,
 * B = 1) and is the V_fft of :506.  spec is complex64 [B][n_fft/2+1][C] contiguous. */
int spectre_rfft_fwd(const void *v, int v_dtype, int64_t v_stride_b, int64_t v_stride_n,
                     void *spec, int B, int N, int n_fft, int C, void *stream);

/* ---- decode side: one bandwidth-bound pass over the running spectrum prefix_fft, complex64 [n_fft/2+1][d],
 * for all heads of a layer (gate row of channel c = c / group_width).  t is the index of the token being added
 * (PrefixFFTCache.t after its increment); when t >= n_fft the token at ring position t % n_fft is evicted and v_old
 * must hold it.  Phase angles are evaluated in float32 in the reference's rounding order.
 *
 * spectre_decode_update   replaces PrefixFFTCache.decode_step's spectrum update, spectre.py:795-806
 * spectre_decode_readout  replaces `gate_broadcast * prefix_fft` + pruned_irfft_single, spectre.py:605, :614-655;
 *                         gate = complex64 [d/group_width][n_fft/2+1] (already multiplied by the positional phase of
 *                         spectre.py:594-598), out = float32 [d], pos = output sample index
 * spectre_decode_step     both in one pass: update with (v_new, v_old, t), then read sample t % n_fft out
 * The read-out reduces over frequency in two stages (per-block partial sums in `workspace`, then a fixed-order sum), so a
 * token is bit-reproducible run to run like the reference's `sum(dim=0)`; workspace = device memory of
 * spectre_decode_workspace_bytes(n_fft, d) bytes, owned by the caller (one per cache). */
size_t spectre_decode_workspace_bytes(int n_fft, int d);
int spectre_decode_update(void *prefix_fft, const float *v_new, const float *v_old, int n_fft, int d, long long t, void *stream);
int spectre_decode_readout(const void *prefix_fft, const void *gate, float *out, int n_fft, int d, int group_width, int pos,
                           void *workspace, size_t workspace_bytes, void *stream);
int spectre_decode_step(void *prefix_fft, const float *v_new, const float *v_old, const void *gate, float *out, int n_fft,
                        int d, int group_width, long long t, void *workspace, size_t workspace_bytes, void *stream);

/* Gate of SpectreHead.decode_step (spectre.py:586-598) in ONE launch: cubic interpolation of the anchors, modReLU and the
 * decode-time positional phase exp(j * 2 pi k (t - t % n_fft) / n_fft), the angle evaluated in float32 in the reference's
 * rounding order.  anchors complex64 [NG][Bk] (gate_mlp output of the running descriptor), bias [NG][F_half], eps [NG],
 * gate complex64 [NG][F_half]; G = gate rows per head as in spectre_gate_expand. */
int spectre_decode_gate(const void *anchors, const float *bias, const float *eps, void *gate, int NG, int G, int Bk,
                        int F_half, long long t, int n_fft, void *stream);

/* ---- gate generator tail: for ALL heads of a layer in one launch,
 *   gate[b, g, k] = modReLU_g( cubic_interp(planes of anchors[b, head(g)])(k) ) * pos_phase[b or 0, k]
 * where, as in the reference (spectre.py:41 reshapes a (B, 2, G, K) stack to (B*G, 2, 1, K)), the real / imaginary
 * planes of gate row j of a head are rows 2j and 2j+1 of the list [re_0 .. re_{G-1}, im_0 .. im_{G-1}] of that head's
 * anchor rows; G = gate rows per head (NG = heads * G).
 * Replaces, per head, interp_complex_1d(..., mode="cubic") spectre.py:526-528 (-> :26-61: grid_sample bicubic, border,
 * align_corners=True), ComplexModReLU spectre.py:531 (-> :109-121) and the positional phase spectre.py:534-536.
 *   anchors    device, complex64 [B][NG][Bk]   (gate_mlp output viewed as complex, spectre.py:515-516, heads stacked)
 *   bias       device, float32 [NG][F_half]    (modrelu.bias of every head, stacked)
 *   eps        device, float32 [NG]            (modrelu.eps of the head each gate row belongs to)
 *   pos_phase  device or NULL, complex64 [*][F_half]; pos_stride_b = F_half for a per-sample phase, 0 for a shared one
 *   gate       device, complex64 [B][NG][F_half] */
int spectre_gate_expand(const void *anchors, const float *bias, const float *eps, const void *pos_phase,
                        long long pos_stride_b, void *gate, int B, int NG, int G, int Bk, int F_half, void *stream);

/* Tuning / introspection used by bench.py and the tests (not needed by a binding). */
typedef struct spectre_mix_plan_info {
    int n_fft;
    int radix[4];          /* stage radices, product = n_fft, unused = 1 */
    int tile_channels;     /* channels per CTA tile */
    int threads;           /* threads per CTA */
    int ctas_per_sm;       /* resident CTAs per SM the launch is sized for */
    int smem_bytes;        /* dynamic shared memory per CTA */
    int grid;              /* CTAs launched for the given problem */
    int launches;          /* kernel launches per call */
    int64_t algorithmic_bytes; /* V + out + gate (+ mem) bytes of the call */
    int64_t workspace_bytes;   /* = spectre_mix_workspace_bytes(...) */
    int dit;               /* 2: the transform runs as two interleaved half-length sub-transforms (even / odd rows) combined by a
                              radix-2 butterfly inside the kernel's middle pass (n_fft = 8192 fp32); radix[] then describes the
                              sub-transform.  1 otherwise */
} spectre_mix_plan_info;

int spectre_mix_plan(int v_dtype, int out_dtype, int has_mem, int B, int N, int n_fft, int C,
                     int group_width, spectre_mix_plan_info *info);

/* Override the tile width (channels per CTA; 0 = automatic).  For experiments only. */
int spectre_mix_set_tile_channels(int tile_channels);

/* L2 prefetch of a CTA's next tiles: 0 off (default), 1 TMA prefetch, 2 cooperative whole-line prefetch (diagnostic build).  For experiments only. */
int spectre_mix_set_prefetch(int enable);

/* Enable (default) / disable TMA-staged tile loads (falls back to direct 128-bit global loads).  For experiments only. */
int spectre_mix_set_tma(int enable);

/* Enable (default) / disable staging of tile I/O through tensor memory (TMEM) by a helper warpgroup, where a variant
 * exists (n_fft = 4096 fp32).  For experiments only. */
int spectre_mix_set_tmem(int enable);

/* Long transforms (n_fft > 4096): 0 = single-kernel variants only, 1 = automatic (default: a single TMEM-staged kernel where
 * one exists -- n_fft = 8192 fp32 -- else the three-launch path around a workspace), 2 = the three-launch path wherever possible.
 * spectre_mix_set_sched bit 5 (32) switches the DIT2 variant of n_fft = 8192 off (experiments: falls back to Plan<16,2,16,16>). */
int spectre_mix_set_two_pass(int enable);

/* Warp stagger before the warp-local passes: 0 = off; code > 0: hold back half of the warps by `code` nanoseconds;
 * code < 0: hold back the warps of scheduler slot s by s * (|code| % 100000) clock cycles (|code| / 100000 picks the
 * grouping: 0 four steps, 1 slots {0,1} vs {2,3}, 2 even vs odd). */
int spectre_mix_set_skew_ns(int code);

/* L2 promotion of the input tensor map (TMA loads): 0 none (default), 1 = 64 B, 2 = 128 B, 3 = 256 B.  For experiments. */
int spectre_mix_set_l2_promotion(int level);

/* Scheduling flags (default 3): bit 0 stagger also before the last inverse pass; bit 1 split barrier around the last inverse
 * pass's shared-memory read (n_fft = 4096 kernel); bit 6 (64) applies the warp stagger to the variants without tensor-memory staging too (experiments); bit 7 (128) switches programmatic dependent launch OFF.  By default every mix
 * launch carries cudaLaunchAttributeProgrammaticStreamSerialization: the kernel's set-up (barriers, tensor-memory allocation,
 * twiddle table) may overlap the tail of the previous kernel of the stream, and it executes griddepcontrol.wait before it reads
 * or writes any tensor, so stream order of all data accesses is unchanged (also under CUDA-graph capture). */
int spectre_mix_set_sched(int flags);

/* Debug: device buffer of grid * 5 groups (4 compute thread groups + the helper warpgroup) * 8 tiles * 8 uint64 that receives per-phase %globaltimer stamps
 * (NULL = off). */
int spectre_mix_set_timeline(void *device_buffer);

#ifdef __cplusplus
}
#endif
#endif /* SPECTRE_MIX_H_ */
