/* iss_b200.h -- C ABI of libiss_b200.so: the B200 (sm_100a) implementation of
 * inaSpeechSegmenter's per-frame hot path.
 *
 * The reference (ina-foss/inaSpeechSegmenter) is pure Python and has no FFI;
 * its seams are Python call signatures.  Every entry point below names the
 * reference call it replaces (path:line relative to the reference tree).
 * Conventions:
 *   - plain C types only; all `d_*` pointers are DEVICE pointers owned by the
 *     caller (e.g. torch.Tensor.data_ptr()), `h_*` pointers are HOST pointers;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     calls are asynchronous on that stream unless stated otherwise;
 *   - return value: 0 = ISS_OK, negative = error; iss_last_error() returns a
 *     thread-local message for the last failing call on this thread;
 *   - handles are opaque; distinct contexts may be used concurrently from
 *     different host threads (the reference extracts features on a worker
 *     thread while the main thread runs the CNNs, segmenter.py:377-387);
 *   - there is NO CPU fallback anywhere behind this ABI.
 */
#ifndef ISS_B200_H
#define ISS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISS_OK              0
#define ISS_ERR_INVALID    -1   /* bad argument */
#define ISS_ERR_CUDA       -2   /* a CUDA runtime call failed */
#define ISS_ERR_NOMEM      -3
#define ISS_ERR_UNSUPPORTED -4  /* e.g. layer type / shape the kernels do not cover */
#define ISS_ERR_STATE      -5   /* tables not uploaded, model not loaded, ... */

typedef struct iss_ctx iss_ctx;
typedef struct iss_cnn iss_cnn;
typedef struct iss_resnet iss_resnet;
typedef struct iss_mlp iss_mlp;

/* ---- library / context ------------------------------------------------- */

#define ISS_ABI_VERSION 2
int         iss_version(void);               /* == ISS_ABI_VERSION of the header the library was built from */
const char *iss_last_error(void);
/* Creates a context bound to CUDA device `device` (one per host thread that
 * issues work).  Fails with ISS_ERR_CUDA if the device is not compute
 * capability 10.x. */
int iss_ctx_create(int device, iss_ctx **out);
int iss_ctx_destroy(iss_ctx *ctx);
/* Number of kernels this library has launched in this process so far (all
 * contexts); bench.py reports the delta over its timed region. */
int64_t iss_launch_count(void);
/* GEMM engine used by the conv / dense layers of K2 and K5 (process-wide):
 * 0 = fp32 CUDA cores; 2 = tcgen05 3xTF32 (activation operand in tensor memory) for every
 * layer shape; 3 (default) = fp16 hi/lo split on tcgen05 kind::f16 for the un-padded stride-1
 * convolutions of the segmenter CNNs, engine 2 for every other layer.  All are sm_100a code
 * paths of this library with fp32-class accuracy (<= 1e-5 on the softmax against the fp32
 * oracle); the default can be overridden with ISS_B200_GEMM=fp32|tc_ts|tc_f16. */
int iss_set_gemm_mode(int mode);
int iss_get_gemm_mode(void);

/* ---- K1: SIDEKIT log-mel + log-energy front-end -------------------------
 * Replaces mfcc(sig, get_mspec=True) as used by _media2feats
 * (inaSpeechSegmenter/sidekit_mfcc.py:278-352 via segmenter.py:53-58):
 * framing 400/160 (:240-263), per-frame pre-emphasis 0.97 (:266-275),
 * loge = log(sum y^2) before windowing (:226), Hann(400) (:223), 512-point
 * real FFT power (:232-233), 24-band mel filterbank (:118-197) and log (:334).
 * The dead DCT (:337) is not computed. */

#define ISS_PCM_F32 0   /* float32 samples in [-1, 1) (what soundfile returns, io.py:51) */
#define ISS_PCM_S16 1   /* int16 PCM; converted as s / 32768.0f, exactly what soundfile does */

#define ISS_FFT_FP32 0  /* single-precision FFT (fast) */
#define ISS_FFT_FP64 1  /* double-precision window + FFT, power rounded to f32: the
                           reference's own precision recipe (sidekit_mfcc.py:231-233) */

/* L = int((n - 400) / 160) + 1 frames, 0 if n < 400 (sidekit_mfcc.py:254). */
int64_t iss_sidekit_num_frames(int64_t n_samples);

/* Host-precomputed tables: fbank = trfbank(16000,512,100,8000,0,24)[0]
 * (float32 [24][257], sidekit_mfcc.py:332) and window = numpy.hanning(400)
 * (float64 [400], :223).  Must be called once per context before
 * iss_sidekit_features.  Synchronous. */
int iss_sidekit_upload_tables(iss_ctx *ctx, const float *h_fbank, const double *h_window);

/* d_pcm: n_samples samples (format pcm_format) in device memory.
 * d_mspec: float32 [L][24]; d_loge: float32 [L];
 * d_loge_stats: double[2] = { sum of finite loge, count of finite loge } --
 * the global reduction _energy_activity needs (segmenter.py:70); written by a
 * deterministic two-stage reduction (no atomics).  A time-shard passes a
 * pointer to its own first sample (frame f starts at sample 160*f); int16
 * input needs 2-byte and float input 4-byte alignment only. */
int iss_sidekit_features(iss_ctx *ctx, const void *d_pcm, int pcm_format, int64_t n_samples,
                         int fft_precision, float *d_mspec, float *d_loge,
                         double *d_loge_stats, void *stream);

/* Stand-alone version of the {sum, count} reduction over a float32 loge array
 * (e.g. the all-gathered loge of a time-sharded recording): same summation
 * order as the fused reduction inside iss_sidekit_features, hence bit-identical
 * d_loge_stats for the same loge values. */
int iss_loge_stats(iss_ctx *ctx, const float *d_loge, int64_t L, double *d_loge_stats, void *stream);

/* ---- K3: Viterbi smoothing ------------------------------------------------
 * Replaces viterbi_decoding(emission, transition) for the two ways the hot
 * path calls it (inaSpeechSegmenter/pyannote_viterbi.py:118-224; callers
 * segmenter.py:72 and :176).  Exact IEEE-double max-sum recursion in the
 * reference's evaluation order, first-max tie break, uniform prior log(1/K). */

/* Energy activity (segmenter.py:69-73 + viterbi_utils.py:29-42):
 *   thr   = (double)(float)(stats[0]/stats[1]) + log_ratio   (NaN if count == 0)
 *   raw_t = (double)loge[t] > thr
 *   E_t   = { raw_t ? h_emis[1] : h_emis[0],  raw_t ? h_emis[0] : h_emis[1] }
 *           with h_emis = { log(1-1e-10), log(1e-10) } computed by the host the
 *           way the reference does (numpy), A = h_trans (row-major [2][2],
 *           from->to), prior = log_prior = log(1/2).
 * d_states: uint8 [ceil(L / out_stride)] -- state of every out_stride-th frame
 * (the reference keeps [::2], segmenter.py:262).  d_work: scratch of at least
 * iss_viterbi_work_bytes(L, 1) bytes.
 * Tracks of >= 65536 frames are decoded chunk-parallel (16384-frame chunks: max-plus transfer
 * matrices -> entry scores -> concurrent true passes); iss_set_energy_viterbi_serial(1), or
 * ISS_B200_VITERBI=serial, forces the single serial chain (the checker of the GPU tests). */
int iss_set_energy_viterbi_serial(int serial);
int iss_energy_viterbi(iss_ctx *ctx, const float *d_loge, int64_t L, const double *d_loge_stats,
                       double log_ratio, const double *h_emis, const double *h_trans,
                       double log_prior, int out_stride, uint8_t *d_states, void *d_work,
                       void *stream);

/* Batched decode of CNN posteriors (segmenter.py:167-178): for segment s the
 * rows d_probs[seg_off[s] .. seg_off[s+1]) (float32 [n][K], already carrying
 * the 0.5 override of :175) are turned into emissions log(p) (float32 log,
 * widened to double like numpy's promotion) and decoded independently with a
 * fresh uniform prior log_prior = log(1/K).  h_seg_off: n_seg+1 int64 host
 * offsets.  2 <= K <= 4.  d_states: uint8 [n]. */
int iss_viterbi_segments(iss_ctx *ctx, const float *d_probs, int K, const int64_t *h_seg_off,
                         int n_seg, const double *h_trans, double log_prior, uint8_t *d_states,
                         void *d_work, void *stream);
int64_t iss_viterbi_work_bytes(int64_t total_steps, int n_seg);

/* Partial chains of the energy Viterbi for a recording that is time-sharded across GPUs
 * (the whole-file chain of pyannote_viterbi.py:202-220 cut at rank boundaries):
 *  iss_energy_transfer: max-plus transfer matrix of this rank's L frames, h_matrix[j*2+i] = best
 *      score of ending in state j given the chain entered (before the first frame) in state i --
 *      two basis-vector forward chains, no back-pointers.  Synchronous.
 *  iss_energy_forward: the true forward pass over the L frames, from the sequence start
 *      (h_vin == NULL: V[0] = E[0] + log_prior) or continuing from the predecessor's outgoing
 *      scores h_vin[2]; stores back-pointers in d_work, returns the outgoing scores h_vout[2] and
 *      the composite back map h_backmap[x] = state just before the first frame given state x at
 *      the last frame.  Synchronous.
 *  iss_energy_emit: backtrack from end_state (-1: argmax of the outgoing scores of the preceding
 *      iss_energy_forward on the same d_work) and write every out_stride-th state.
 * d_work: iss_viterbi_work_bytes(L, 1) bytes, the same buffer for forward and emit. */
int iss_energy_transfer(iss_ctx *ctx, const float *d_loge, int64_t L, const double *d_loge_stats,
                        double log_ratio, const double *h_emis, const double *h_trans,
                        double *h_matrix, void *d_work, void *stream);
int iss_energy_forward(iss_ctx *ctx, const float *d_loge, int64_t L, const double *d_loge_stats,
                       double log_ratio, const double *h_emis, const double *h_trans, double log_prior,
                       const double *h_vin, double *h_vout, uint8_t *h_backmap, void *d_work, void *stream);
int iss_energy_emit(iss_ctx *ctx, int64_t L, int end_state, int out_stride, uint8_t *d_states,
                    void *d_work, void *stream);

/* ---- K2: patch z-normalisation + CNN forward ------------------------------
 * Replaces _get_patches (segmenter.py:76-88) + keras Model.predict
 * (segmenter.py:131-133,163) + the non-finite override (:175). */

#define ISS_LAYER_CONV2D  1
#define ISS_LAYER_DENSE   2
#define ISS_LAYER_MAXPOOL 3

#define ISS_F_BIAS        1   /* + bias[cout] */
#define ISS_F_AFFINE_PRE  2   /* then  x*pre_scale[c] + pre_shift[c]   (BatchNorm before the activation) */
#define ISS_F_RELU        4   /* then  max(x, 0) */
#define ISS_F_AFFINE_POST 8   /* then  x*post_scale[c] + post_shift[c] (BatchNorm after the activation) */
#define ISS_F_SOFTMAX     16  /* softmax over cout (last layer only) */
#define ISS_F_SIGMOID     32

typedef struct iss_layer_desc {
    int32_t kind;                 /* ISS_LAYER_* */
    int32_t kh, kw;               /* kernel / pool window (rows = time, cols = mel band) */
    int32_t sh, sw;               /* strides */
    int32_t pad_top, pad_left, pad_bottom, pad_right;  /* zero (conv) / -inf (pool) padding */
    int32_t cin, cout;            /* DENSE: cin = flattened (H*W*C, Keras channels_last order) */
    int32_t flags;                /* ISS_F_* */
    int64_t w_off;                /* float offsets into the weight blob; -1 when unused.     */
    int64_t bias_off;             /* CONV2D kernel is [kh][kw][cin][cout] (Keras layout),     */
    int64_t pre_scale_off, pre_shift_off;    /* DENSE kernel is [cin][cout].                */
    int64_t post_scale_off, post_shift_off;
} iss_layer_desc;

/* Builds a model for input patches of in_h x in_w x 1 (68 x nmel).  The blob is
 * copied to the device.  Synchronous. */
int iss_cnn_create(iss_ctx *ctx, const iss_layer_desc *layers, int n_layers,
                   const float *h_blob, int64_t blob_len, int in_h, int in_w, iss_cnn **out);
int iss_cnn_destroy(iss_cnn *cnn);
int iss_cnn_num_classes(const iss_cnn *cnn);
double iss_cnn_flops_per_patch(const iss_cnn *cnn);     /* 2 * MACs of all conv/dense layers */
/* scratch needed by iss_cnn_forward for `n` patches in `n_seg` index ranges
 * (activations ping-pong, patch statistics, index map). */
int64_t iss_cnn_workspace_bytes(const iss_cnn *cnn, int64_t n, int n_seg);

/* d_mspec: float32 [L][ld] log-mel rows (ld = 24; the first in_w bands are
 * used, segmenter.py:146-147).  Patch p (0 <= p < ceil(L/2)) is frames
 * 2*clamp(p-17, 0, U-1) .. +67 with U = (L-68)/2+1 un-replicated windows
 * (edge replication of segmenter.py:83-85); edge_left/edge_right = 0 disable
 * the replication on that side for a time-shard that is not at a file end
 * (then patch p is frames 2p .. 2p+67 and p < U).
 * The patches evaluated are those of the index ranges
 * [h_seg_start[s], h_seg_stop[s]) concatenated (segmenter.py:156-162).
 * d_probs: float32 [n][K] softmax, rows of non-finite patches forced to 0.5
 * (segmenter.py:175).  n = sum of range lengths. */
int iss_cnn_forward(iss_ctx *ctx, iss_cnn *cnn, const float *d_mspec, int64_t L, int ld,
                    int edge_left, int edge_right,
                    const int32_t *h_seg_start, const int32_t *h_seg_stop, int n_seg,
                    float *d_probs, void *d_work, int64_t work_bytes, void *stream);

/* ---- K4: VBx / HTK 64-band log-mel front-end + floating CMVN ----------------
 * Replaces get_features(signal) (inaSpeechSegmenter/vbx_segmenter.py:72-89 on top
 * of features_vbx.py:62-148): trunc(signal*2^15) + dither, mirror pad 120/200,
 * frames 400/160, per-frame mean removal + pre-emphasis, Povey window, 512-point
 * power spectrum, log(max(1, . fbank)), cmvn_floating_kaldi(150, 149), float32. */

/* M = (n + 320 - 400) / 160 + 1 feature frames (0 if n < 200). */
int64_t iss_vbx_num_frames(int64_t n_samples);
/* h_fbank = mel_fbank_mx(400, 16000, NUMCHANS=64, LOFREQ=20, HIFREQ=7600,
 * htk_bug=False) (float64 [257][64]); h_window = povey_window(400) (float64). */
int iss_vbx_upload_tables(iss_ctx *ctx, const double *h_fbank, const double *h_window);
int64_t iss_vbx_work_bytes(int64_t n_samples);
/* d_dither: float64 [n] = 8*(2u-1), u the np.random.seed(3); np.random.rand(n)
 * stream (a prefix of one fixed MT19937 sequence -- keep it resident), or NULL for
 * no dither.  d_fea: float32 [M][64].  d_work: iss_vbx_work_bytes(n) of scratch. */
int iss_vbx_features(iss_ctx *ctx, const void *d_pcm, int pcm_format, int64_t n_samples,
                     const double *d_dither, float *d_fea, void *d_work, void *stream);

/* ---- K5: ResNet101 x-vector extractor ----------------------------------------
 * Replaces VBxExtractor.get_embedding / OnnxBackendExtractor (vbx_segmenter.py:
 * 249-266) == ResNet.forward (resnet.py:115-130), batched over the windows of
 * VBxExtractor.__call__ (vbx_segmenter.py:217-246).
 * Blob layout (float32), in module order conv1, layer1.0 ... layer4.2, each
 * Bottleneck as conv1, conv2, conv3[, shortcut]; every convolution contributes
 * kernel [kh][kw][cin][cout], then BatchNorm folded to scale[cout], shift[cout];
 * finally embedding weight transposed [2*C*H][embed] and bias [embed]. */
int64_t iss_resnet_blob_len(int m_channels, int feat_dim, int embed_dim, const int *num_blocks);
int iss_resnet_create(iss_ctx *ctx, const float *h_blob, int64_t blob_len, int m_channels, int feat_dim,
                      int embed_dim, const int *num_blocks, iss_resnet **out);
int iss_resnet_destroy(iss_resnet *net);
double iss_resnet_flops_per_window(const iss_resnet *net, int win_len);
int64_t iss_resnet_workspace_bytes(const iss_resnet *net, int n_windows, int win_len);
/* d_fea: float32 [M][feat_dim] (CMVN'd features); window i = rows
 * [h_win_start[i], h_win_start[i] + win_len); d_emb: float32 [n_windows][embed]. */
int iss_resnet_embed(iss_ctx *ctx, iss_resnet *net, const float *d_fea, int64_t M,
                     const int32_t *h_win_start, int n_windows, int win_len, float *d_emb,
                     void *d_work, int64_t work_bytes, void *stream);

/* ---- Dense stack applied to x-vectors ------------------------------------------------
 * Replaces gender_detection_mlp_model.predict(x) of VoiceFemininityScoring
 * (inaSpeechSegmenter/vbx_segmenter.py:116-124,188-191).  layers: ISS_LAYER_DENSE records
 * only (flags as for the CNN head, ISS_F_SIGMOID supported, no softmax). */
int iss_mlp_create(iss_ctx *ctx, const iss_layer_desc *layers, int n_layers, const float *h_blob,
                   int64_t blob_len, int in_dim, iss_mlp **out);
int iss_mlp_destroy(iss_mlp *mlp);
int iss_mlp_out_dim(const iss_mlp *mlp);
int64_t iss_mlp_workspace_bytes(const iss_mlp *mlp, int64_t n_rows);
/* d_x: float32 [n_rows][in_dim]; d_y: float32 [n_rows][out_dim]. */
int iss_mlp_forward(iss_ctx *ctx, iss_mlp *mlp, const float *d_x, int64_t n_rows, float *d_y,
                    void *d_work, int64_t work_bytes, void *stream);

/* Live roofline support: record CUDA events around every launch of layer
 * `layer` (index into the iss_layer_desc list; -1 switches profiling off) on
 * the stream the launch uses.  iss_cnn_profile_read synchronises those events,
 * returns the summed device time (ms), the number of launches and the FLOPs
 * those launches performed, and resets the accumulators. */
int iss_cnn_profile(iss_cnn *cnn, int layer);
int iss_cnn_profile_read(iss_cnn *cnn, double *total_ms, int64_t *launches, double *flops);
/* FLOPs of one layer for one patch (2 * MACs); 0 for pooling. */
double iss_cnn_layer_flops(const iss_cnn *cnn, int layer);
int iss_cnn_num_layers(const iss_cnn *cnn);

#ifdef __cplusplus
}
#endif
#endif /* ISS_B200_H */
