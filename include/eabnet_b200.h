/*
 * eabnet_b200 -- C ABI of the B200-native EaBNet inference hot path.
 *
 * The reference (Ezreal11/EaBNet) is pure Python/PyTorch and has no FFI of its own; the boundary it exposes
 * for this path is the Python class `EaBNet(nn.Module)` (EaBNet.py:9-125).  Every entry point below names
 * the reference interface it stands in for.  Signatures use plain pointers and sizes only (no torch types);
 * "dev" pointers are CUDA device pointers on the current device, `stream` is a cudaStream_t passed as
 * void* (NULL = legacy default stream).  All functions return 0 on success, non-zero on failure with a
 * message available from eab_last_error() (the Python host layer turns that into RuntimeError -- the
 * reference reports errors as Python exceptions only).  There is no CPU fallback.
 *
 * Tensor layouts are the reference's own at the boundary:
 *     inpt  [B, T, F, M, 2] fp32   (EaBNet.forward argument, EaBNet.py:90)
 *     out   [B, 2, T, F]    fp32   (EaBNet.forward result,   EaBNet.py:91);  [B, 2, T] for topo_type "miso"
 *     wave  [B, M, L]       fp32   (prepare_data argument,   test.py:20,32)
 *     enhanced wave [B, 160*(L/160)] fp32 (torch.istft result, enhance.py:61-62)
 */
#ifndef EABNET_B200_H_
#define EABNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define EAB_API __attribute__((visibility("default")))
#else
#define EAB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* Constructor arguments of EaBNet.__init__ (EaBNet.py:10-27), same meaning, enums for the strings. */
typedef struct eab_config {
    int k1_t, k1_f;      /* k1 = (2,3)  kernel of the gated 2-D convs                       */
    int k2_t, k2_f;      /* k2 = (1,3)  kernel inside the inner U-Nets                      */
    int c;               /* 64          channels of the 2-D convs                           */
    int M;               /* 9           microphones                                        */
    int embed_dim;       /* 64                                                             */
    int kd1;             /* 5           dilated kernel size in the squeezed TCM             */
    int cd1;             /* 64          squeezed channels                                  */
    int d_feat;          /* 256         TCM feature channels (= 64 * bottleneck F)          */
    int p, q;            /* 6, 3        TCMs per group, groups                              */
    int is_causal;       /* 1                                                              */
    int is_u2;           /* 1                                                              */
    int bf_type;         /* 0 "lstm", 1 "cnn"                                               */
    int topo_type;       /* 0 "mimo", 1 "miso"                                              */
    int intra_connect;   /* 0 "cat",  1 "add"                                               */
    int norm_type;       /* 0 "IN",   1 "BN" (eval-mode running statistics)                 */
    int n_freq;          /* 161         F of the input spectrum (fft_num/2+1, test.py:26)   */
} eab_config;

typedef struct eab_model eab_model;

/* kinds reported by eab_param_info (used by the host layer to pick torch's default initialiser) */
enum { EAB_P_CONV_W = 0, EAB_P_CONV_B = 1, EAB_P_NORM_G = 2, EAB_P_NORM_B = 3, EAB_P_PRELU = 4,
       EAB_P_LSTM = 5, EAB_P_LIN_W = 6, EAB_P_LIN_B = 7, EAB_P_BN_MEAN = 8, EAB_P_BN_VAR = 9,
       EAB_P_BN_COUNT = 10 };

/* EaBNet.__init__ (EaBNet.py:50-86).  Pure host work: no CUDA call is made until parameters are
 * committed, so this (and the parameter table below) is usable on a machine without a GPU. */
EAB_API int  eab_create(const eab_config* cfg, eab_model** out);
EAB_API void eab_destroy(eab_model* m);

/* nn.Module.state_dict() key/shape table (498 entries for the default config), in the
 * reference's registration order.  `shape` receives up to 4 extents, `fan_in` the torch fan-in of the owning
 * layer (for default initialisation). */
EAB_API int  eab_param_count(const eab_model* m);
EAB_API int  eab_param_info(const eab_model* m, int index, const char** name, int* ndim, int64_t shape[4],
                    int* kind, int* fan_in);

/* nn.Module.load_state_dict (enhance.py:22, test.py:165): hand over one fp32 tensor in its reference layout
 * from HOST memory; eab_commit_params packs everything into the kernels' layouts and uploads it (one blob).
 * BN "num_batches_tracked" is accepted and ignored. */
EAB_API int  eab_set_param(eab_model* m, const char* name, const float* host_data, int64_t numel);
EAB_API int  eab_commit_params(eab_model* m, void* stream);

/* Scratch needed by eab_forward / eab_enhance for a given batch and frame count (caller allocates; the
 * reference relies on torch's caching allocator, SURVEY.md section 8b "Ownership"). */
EAB_API size_t eab_workspace_bytes(const eab_model* m, int B, int T);

/* EaBNet.forward (EaBNet.py:88-125).  inpt is not modified. */
EAB_API int  eab_forward(eab_model* m, const float* inpt_dev, float* out_dev, int B, int T,
                 void* workspace_dev, size_t workspace_bytes, void* stream);

/* prepare_data's noisy branch (test.py:32-43 == train_distributed.py:80-91): STFT 320/160 hann,
 * center/reflect, + square-root magnitude compression.  T = 1 + L/160.  spec_dev [B,T,161,M,2]. */
EAB_API int  eab_stft(const float* wave_dev, float* spec_dev, int B, int M, int L, void* stream);

/* torch.istft call of enhance.py:59-61 / test.py:189-190 on a [B,2,T,161] spectrum -> [B,160*(T-1)]. */
EAB_API int  eab_istft(const float* spec_dev, float* wave_dev, int B, int T, void* stream);

/* The whole test.py:178-190 sequence on device buffers: wave [B,M,L] -> enhanced [B,160*(L/160)].
 * Workspace must be at least eab_enhance_workspace_bytes(m,B,L). */
EAB_API size_t eab_enhance_workspace_bytes(const eab_model* m, int B, int L);
EAB_API int  eab_enhance(eab_model* m, const float* wave_dev, float* enhanced_dev, int B, int L,
                 void* workspace_dev, size_t workspace_bytes, void* stream);

/* Same with HOST buffers (the call enhance.py makes end to end: H2D, compute, D2H all inside; the library
 * keeps its own device scratch).  Synchronises the stream before returning. */
EAB_API int  eab_enhance_host(eab_model* m, const float* wave_host, float* enhanced_host, int B, int L, void* stream);

/* Dataset-scale form of the same call (enhance.py's loop over files; BASELINE configs[3]): n_batches batches of B
 * utterances, waves_host[i] -> [B,M,L], enhanced_host[i] -> [B,160*(L/160)], HOST (ideally pinned) buffers.  The
 * upload of batch i+1 and the download of batch i-1 run on the library's own copy streams while batch i computes on
 * `stream` (double-buffered device slots).  Synchronises before returning. */
EAB_API int  eab_enhance_host_batches(eab_model* m, const float* const* waves_host, float* const* enhanced_host,
                              int n_batches, int B, int L, void* stream);

/* Causal frame-by-frame inference with carried state (BASELINE configs[2]; the reference has no streaming entry point --
 * the semantics are those of EaBNet.forward restricted to is_causal=True (EaBNet.py:46-48): frame n of the output
 * depends on input frames <= n only, so n_streams concurrent streams stepped one 10 ms hop at a time reproduce the
 * offline result frame for frame).  Needs norm_type "BN" (InstanceNorm statistics span the whole utterance).
 * The caller owns the state blob (device memory, 256-byte aligned, eab_stream_state_bytes(m, n_streams) bytes): conv
 * history rings, TCM dilation rings, LSTM (h, c), the previous hop and the overlap-add tail.  One step launches
 * kernels only (the frame counter lives in the state and is read on the device): it can be captured in a CUDA graph.
 * Kernel-selection options ("umma", "stream_umma", "stream_lstm", "stream_tcm") change the layout too: set them before
 * eab_stream_state_bytes / eab_stream_reset; after changing one, size and reset the blob again.
 * Size the blob AFTER eab_commit_params: which layers run on the tensor cores (padded channel counts of their outputs) is
 * known once the weights are packed; a blob sized before that is refused by reset / step with a message, never overrun.
 *   eab_stream_step      hop_dev [S][M][160] new samples -> enhanced_hop_dev [S][160], delayed by ONE hop (overlap-add
 *                        needs the next frame): call k returns samples [160(k-1), 160k) of torch.istft's output
 *                        (enhance.py:59-62); the first call returns zeros.
 *   eab_stream_step_spec frame_dev [S][F][M][2] (one column of prepare_data's output) -> out_frame_dev [S][2][F]. */
EAB_API size_t eab_stream_state_bytes(const eab_model* m, int n_streams);
EAB_API int  eab_stream_reset(eab_model* m, void* state_dev, size_t state_bytes, int n_streams, void* stream);
EAB_API int  eab_stream_step(eab_model* m, void* state_dev, size_t state_bytes, const float* hop_dev,
                     float* enhanced_hop_dev, int n_streams, void* stream);
EAB_API int  eab_stream_step_spec(eab_model* m, void* state_dev, size_t state_bytes, const float* frame_dev,
                          float* out_frame_dev, int n_streams, void* stream);
/* The servable front door around the step:
 *   eab_stream_step_pcm16  the same step on the 16-bit PCM wire format: hop_dev [S][M][160] int16 (sample / 32768, what
 *                          torchaudio.load yields, enhance.py:35) -> enhanced_hop_dev [S][160] int16(clip(y,-1,1) * 32767)
 *                          (dataset/mcse_dataset_offline_gen.py:38-39); the conversions live in the STFT / iSTFT frame kernels.
 *   eab_stream_reset_one   stream `idx` leaves and a new one joins in its slot: from the next step on it starts over at its frame
 *                          0 (zero causal history, zero LSTM state, reflected first half-frame) while the other streams carry on
 *                          bit-identically.  Stream-ordered (capturable), no synchronisation. */
EAB_API int  eab_stream_step_pcm16(eab_model* m, void* state_dev, size_t state_bytes, const int16_t* hop_dev,
                           int16_t* enhanced_hop_dev, int n_streams, void* stream);
EAB_API int  eab_stream_reset_one(eab_model* m, void* state_dev, size_t state_bytes, int n_streams, int idx, void* stream);

/* Post-filter streaming (enhance.py:49-62 as a causal stream; GaGNet.py:75-89 one frame at a time).  A GaGNet handle built with
 * is_causal and norm_type "BN" takes the same state calls as an EaBNet handle: eab_stream_state_bytes / eab_stream_reset /
 * eab_stream_reset_one (its state holds a frame counter, the per-stream start frames, the q estimates of the step and every
 * activation ring).
 *   eab_gag_stream_step_spec   GaGNet.forward(inpt, pre_x) for frame n of every stream: inpt_frame / pre_frame [S][2][F] ->
 *                              out_frames [q][S][2][F] (every module's estimate of the frame)
 *   eab_stream_step_postnet    one 10 ms hop through EaBNet AND the post-filter: STFT frame -> EaBNet step -> GaGNet step on
 *                              (microphone ref_mic of the compressed spectrum, EaBNet's estimate) -> iSTFT of the last module's
 *                              estimate.  Both states advance together; reset them together.  _pcm16: int16 hops in and out. */
EAB_API int  eab_gag_stream_step_spec(eab_model* gagnet, void* gag_state_dev, size_t gag_state_bytes, const float* inpt_frame_dev,
                              const float* pre_frame_dev, float* out_frames_dev, int n_streams, void* stream);
EAB_API int  eab_stream_step_postnet(eab_model* eabnet, void* state_dev, size_t state_bytes, eab_model* gagnet, void* gag_state_dev,
                             size_t gag_state_bytes, int ref_mic, const float* hop_dev, float* enhanced_hop_dev, int n_streams,
                             void* stream);
EAB_API int  eab_stream_step_postnet_pcm16(eab_model* eabnet, void* state_dev, size_t state_bytes, eab_model* gagnet,
                                   void* gag_state_dev, size_t gag_state_bytes, int ref_mic, const int16_t* hop_dev,
                                   int16_t* enhanced_hop_dev, int n_streams, void* stream);

/* Streaming an InstanceNorm-trained model.  InstanceNorm statistics span the whole utterance (EaBNet.py:684-686), so an IN
 * model cannot be stepped causally; the reference's own note (EaBNet.py:45-48) points at accumulated statistics or another norm.
 * What this library offers is the second: the per-channel statistics every InstanceNorm layer saw on calibration audio become
 * the running_mean / running_var of the same network built with norm_type = "BN" (same weights, same state_dict the reference's
 * EaBNet(norm_type="BN") loads), which streams.  On the calibration utterance itself (B = 1) the two models are the same function.
 *   option "norm_log" = 1, then eab_forward: records every InstanceNorm the forward applied (use "tcm_chain" = 0 as well: the
 *   chain kernel keeps the TCM statistics to itself);  eab_norm_stats_count / eab_norm_stats read them back:
 *   weight_name = the layer's "....norm.weight" parameter, sums_host [C][2] = (sum, sum of squares) over `count` positions. */
EAB_API int  eab_norm_stats_count(const eab_model* m);
EAB_API int  eab_norm_stats(eab_model* m, int i, const char** weight_name, int* C, int64_t* count, double* sums_host, void* stream);


/* ---------------------------------------------------------------------------------------------------------------
 * First slice of the training step (train_distributed.py:214-230, BASELINE configs[4]): the tail
 * of the beamforming head and the loss, forward AND backward, as hand-written kernels (fp32 CUDA cores; no autograd graph, no
 * library call).  Everything else of the backward pass is not built: the module still refuses autograd (model.py).
 *   eab_head_forward    out [B,2,T,F] = sum_m w_m x_m with w = Linear(ReLU(Linear(h2)))   (LSTM_BF.w_dnn, EaBNet.py:593-597,
 *                       612-613, and the filter-and-sum of EaBNet.forward, EaBNet.py:114-117).  W1 [64,64], b1 [64], W2 [2M,64],
 *                       b2 [2M] in torch's layouts; h2 [B,T,F,64] (the second LSTM's output, channels last); spec [B,T,F,M,2].
 *   eab_head_backward   given d_out [B,2,T,F]: d_h2 [B,T,F,64] and grads = [dW1 64x64 | db1 64 | dW2 2Mx64 | db2 2M]
 *                       (summed in a fixed order: bit-identical from run to run); workspace of eab_head_backward_workspace_bytes.
 *   eab_loss_com_mag_mse / _backward   com_mag_mse_loss(esti, label, frame_list) (EaBNet.py:627-640) and its gradient with
 *                       respect to esti; frames_dev [B] (NULL = every utterance has T frames, train_distributed.py:221),
 *                       frames_total = their sum; scratch16 = 16 bytes of device scratch; grad_loss_dev = upstream gradient of
 *                       the scalar (NULL = 1). */
EAB_API int    eab_head_forward(const float* W1_dev, const float* b1_dev, const float* W2_dev, const float* b2_dev, const float* h2_dev,
                                const float* spec_dev, float* out_dev, int B, int T, int F, int M, void* stream);
EAB_API size_t eab_head_backward_workspace_bytes(int M);
EAB_API int    eab_head_backward(const float* W1_dev, const float* b1_dev, const float* W2_dev, const float* b2_dev,
                                 const float* h2_dev, const float* spec_dev, const float* d_out_dev, float* d_h2_dev, float* grads_dev,
                                 int B, int T, int F, int M, void* workspace_dev, size_t workspace_bytes, void* stream);
EAB_API int    eab_loss_com_mag_mse(const float* esti_dev, const float* label_dev, const int* frames_dev, int64_t frames_total, int B,
                                    int T, int F, float* loss_dev, void* scratch16_dev, void* stream);
EAB_API int    eab_loss_com_mag_mse_backward(const float* esti_dev, const float* label_dev, const int* frames_dev, int64_t frames_total,
                                             int B, int T, int F, const float* grad_loss_dev, float* d_esti_dev, void* stream);
/* ... and on frequency-major tensors [B,2,F,T]: one stage of stagewise_com_mag_mse_loss (GaGNet.py:601-619; the stage weights
 * 0.1 / 1 and the sum over stages are scalar arithmetic of the caller), i.e. eabnet_with_postnet_loss (EaBNet.py:642-650) is
 * eab_loss_com_mag_mse(esti0) + sum_i alpha_i eab_loss_com_mag_mse_fm(esti1_list[i]). */
EAB_API int    eab_loss_com_mag_mse_fm(const float* esti_dev, const float* label_dev, const int* frames_dev, int64_t frames_total, int B,
                                       int T, int F, float* loss_dev, void* scratch16_dev, void* stream);
EAB_API int    eab_loss_com_mag_mse_fm_backward(const float* esti_dev, const float* label_dev, const int* frames_dev,
                                                int64_t frames_total, int B, int T, int F, const float* grad_loss_dev,
                                                float* d_esti_dev, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * I/O edges of enhance.py: the wav container and the sample-rate conversion.
 *   eab_wav_info / eab_wav_decode   `noisy, sr = torchaudio.load(path)` (enhance.py:35) on the bytes of a RIFF/WAVE file held
 *                                   in HOST memory: planar float32 [channels][frames], integer PCM scaled like torchaudio
 *                                   (uint8: (x-128)/128, int16 / 2^15, int24 / 2^23, int32 / 2^31), float32 / float64 as stored
 *                                   (PCM, IEEE float and WAVE_FORMAT_EXTENSIBLE files).  `planar_pcm16` (optional, 16-bit PCM
 *                                   files only) receives the raw samples for the int16 front doors (eab_enhance_host_pcm16).
 *   eab_resample                    `torchaudio.transforms.Resample(sr, 16000)(noisy)` (enhance.py:36-37) on DEVICE rows
 *                                   [rows][length] -> [rows][eab_resample_length(length, sr, 16000)]: torchaudio's default
 *                                   kernel (sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99, built in float64 and
 *                                   rounded to float32), applied as a polyphase FIR.  orig == new copies (Resample.forward).
 *   eab_wav_encode                  `wavfile.write(path, 16000, esti_wav[0])` (enhance.py:63): scipy's writer byte for byte -
 *                                   a float32 array becomes a WAVE_FORMAT_IEEE_FLOAT file (18-byte fmt chunk + fact chunk),
 *                                   an int16 array (`interleaved_pcm16` non-NULL) a 44-byte-header PCM file.  Samples are
 *                                   interleaved [frames][channels] as in the file.  HOST buffers. */
EAB_API int     eab_wav_info(const void* bytes, size_t n, int* channels, int* sample_rate, int64_t* frames, int* bits_per_sample,
                             int* is_float);
EAB_API int     eab_wav_decode(const void* bytes, size_t n, float* planar, int16_t* planar_pcm16);
EAB_API size_t  eab_wav_encode_bytes(int64_t frames, int channels, int pcm16);
EAB_API int     eab_wav_encode(const float* interleaved, const int16_t* interleaved_pcm16, int64_t frames, int channels,
                               int sample_rate, void* out, size_t capacity);
EAB_API int64_t eab_resample_length(int64_t length, int orig_freq, int new_freq);
EAB_API int     eab_resample(const float* wave_dev, float* out_dev, int rows, int64_t length, int orig_freq, int new_freq,
                             void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * GaGNet post-filter: what `enhance.py` runs behind EaBNet through
 * `EaBNetWithPostNet` (EaBNet.py:127-148, GaGNet.py:5-89).  Constructor arguments of GaGNet.__init__ (GaGNet.py:6-24),
 * same meaning.  The handle is an eab_model: the parameter table / eab_set_param / eab_commit_params / eab_set_option /
 * eab_debug_tap / eab_destroy entry points above apply unchanged (state_dict keys are GaGNet's: `en.*`, `gags.*`). */
typedef struct eab_gag_config {
    int cin;             /* 2     real/imag planes of each of the two inputs (only 2 runs in the reference)   */
    int k1_t, k1_f;      /* (2,3) */
    int k2_t, k2_f;      /* (1,3) */
    int c;               /* 64    */
    int kd1;             /* 3     */
    int cd1;             /* 64    */
    int d_feat;          /* 256   */
    int p;               /* 2     TCM groups per branch                                                      */
    int q;               /* 3     glance-gaze modules                                                        */
    int n_dilas;         /* 4     */
    int dilas[8];        /* 1,2,5,9 */
    int fft_num;         /* 320   */
    int is_u2;           /* 1     */
    int is_causal;       /* 1     */
    int is_squeezed;     /* 0     one TCM stack shared by the real / imaginary residual branches             */
    int acti_type;       /* 0 "sigmoid", 1 "tanh", 2 "relu"  (gain activation, GaGNet.py:165-172)             */
    int intra_connect;   /* 0 "cat", 1 "add" */
    int norm_type;       /* 0 "IN",  1 "BN"  */
} eab_gag_config;

/* GaGNet.__init__ (GaGNet.py:50-73); pure host work like eab_create. */
EAB_API int    eab_gag_create(const eab_gag_config* cfg, eab_model** out);
EAB_API size_t eab_gag_workspace_bytes(const eab_model* m, int B, int T);
/* GaGNet.forward(inpt, pre_x) (GaGNet.py:75-89).
 *   inpt_dev      the [B,2,T,F] input read through explicit element strides (in floats) for its (b, c, t, f) axes, so
 *                 that the reference-microphone view `noisy_stft[..., ref_mic, :]` of a [B,T,F,M,2] spectrum
 *                 (EaBNet.py:141-142) is passed without a copy: strides (T*F*M*2, 1, F*M*2, M*2), pointer offset
 *                 ref_mic*2.  A contiguous [B,2,T,F] tensor has strides (2*T*F, T*F, F, 1).
 *   pre_dev       [B,2,T,F] contiguous (the layout eab_forward writes).
 *   out_dev       [q][B,2,T,F]: every module's estimate, time-major.  The reference returns each as [B,2,F,T]
 *                 (GaGNet.py:133,88); the host layer hands out transposed views, and `esti_stft` of
 *                 EaBNetWithPostNet (EaBNet.py:147) is out_dev[q-1] as it lies. */
EAB_API int    eab_gag_forward(eab_model* m, const float* inpt_dev, const int64_t inpt_strides[4], const float* pre_dev,
                       float* out_dev, int B, int T, void* workspace_dev, size_t workspace_bytes, void* stream);

/* enhance.py:49-62 on device buffers: wave [B,M,L] -> STFT + compression -> EaBNet -> GaGNet on (microphone `ref_mic`,
 * EaBNet's estimate) -> iSTFT of the last module's estimate -> enhanced [B,160*(L/160)]. */
EAB_API size_t eab_enhance_postnet_workspace_bytes(const eab_model* eabnet, const eab_model* gagnet, int B, int L);
EAB_API int    eab_enhance_postnet(eab_model* eabnet, eab_model* gagnet, int ref_mic, const float* wave_dev,
                           float* enhanced_dev, int B, int L, void* workspace_dev, size_t workspace_bytes, void* stream);

/* 16-bit PCM wire format end to end on HOST buffers (enhance.py:35-43 reads the file with
 * torchaudio.load, i.e. int16 / 32768, and permutes the microphones with index_select; dataset/mcse_dataset_offline_gen.py:38-39
 * writes int16(clip(y,-1,1) * 32767)).  pcm_host [B][M][L] int16 in file channel order; microphone m of the model is file
 * channel mic_order[m] (NULL = identity); gagnet may be NULL (EaBNet only); enhanced_pcm_host [B][160*(L/160)] int16.
 * Half the H2D bytes of eab_enhance_host; the library keeps its own device scratch; synchronises before returning. */
EAB_API int    eab_enhance_host_pcm16(eab_model* eabnet, eab_model* gagnet, int ref_mic, const int16_t* pcm_host,
                              const int* mic_order, int16_t* enhanced_pcm_host, int B, int L, void* stream);
/* Dataset-scale form on the PCM wire format: eab_enhance_host_batches with int16 batches in and out (110 MB instead of 221 MB
 * of H2D per 64 x 6 s batch); the conversions are fused into the STFT's operand staging and the iSTFT's store, the uploads /
 * downloads of neighbouring batches overlap compute and every slot's step is replayed from a CUDA graph, exactly as there. */
EAB_API int    eab_enhance_host_batches_pcm16(eab_model* m, const int16_t* const* pcm_host, const int* mic_order,
                                      int16_t* const* enhanced_pcm_host, int n_batches, int B, int L, void* stream);

/* Introspection used by tests and bench: number of kernels launched by the last forward/enhance call, and a
 * copy of a named intermediate of the last eab_forward ("en.0".."en.4", "tcm", "de.0".."de.3", "embed",
 * "h1", "h2", "w") with its normalisation/activation applied, channels-last [B,T,F',C'].  Returns the
 * number of floats written (<= capacity) or -1. */
EAB_API int     eab_last_launch_count(const eab_model* m);
EAB_API int64_t eab_debug_tap(eab_model* m, const char* name, float* dst_dev, int64_t capacity, void* stream);

/* Precision / kernel-selection knobs of the sm_100a path, per model (defaults in parentheses).  Tensor-core operands are fp16:
 * "1 pass" = operands rounded to fp16 (the 10-bit mantissa of TF32), fp32 accumulation; "3 passes" = hi/lo fp16 split of both
 * operands, hi*hi + lo*hi + hi*lo (~22 mantissa bits, fp32-grade).  The splits are unscaled: post-normalisation activations and
 * weights are O(1e-3 .. 1e2); magnitudes below fp16's subnormal step (6e-8) lose their lo part and magnitudes above 65504 would
 * overflow - neither occurs behind a normalisation layer, and the STFT input is audio in [-1, 1].
 *   "umma" (1)          tcgen05 kernels for every eligible layer; 0 = fp32 CUDA-core kernels only
 *   "enc_passes" (3)    encoder 2-D convs incl. their inner U-Nets ("inner_passes" sets the inner U-Nets alone)
 *   "dec_passes" (1)    decoder 2-D convs            "first_passes" (3)  the first gated conv (2M input channels)
 *   "raw" (1)           conv_raw_kernel: the layer's raw fp32 input tiles are normalised in shared memory (no stage pass);
 *                       0 = stage_kernel + conv_tma_kernel (fp16 operand planes staged through HBM)
 *   "staged" (1)        the stage + conv_tma pair for what conv_raw does not take; 0 = per-tap gather kernel (conv_umma)
 *   "lazy" (1)          a module's residual sum x0 + y is summed by its consumers, never materialised
 *   "tcm_chain" (1)     a group of squeezed TCMs as one launch (a thread-block cluster per utterance); 0 = three GEMM launches
 *                       per TCM, 2 = one chain per launch, 3 = the cooperative grid-barrier form (also taken when T > 1024)
 *   "fused_head" (1)    w_dnn + filter-and-sum as one kernel     "head_w_tap" (0)  ... which also writes the beam weights (tap "w")
 *   "stft_tc" (1)       STFT as a tensor-core DFT-GEMM; 0 = fp32 CUDA-core kernel.  PROCESS-WIDE switch (eab_stft has no handle)
 *   "istft_tc" (0)      iSTFT as a tensor-core two-tap DFT-GEMM (window, 1/320 and the overlap-add envelope folded into the weight
 *                       images; 3-pass fp16 split, 2.6e-6) instead of the fused fp32 CUDA-core kernel (4e-8).  Measured slower (0.245
 *                       vs 0.209 ms per 64 x 6 s), hence off.  PROCESS-WIDE like stft_tc
 *   "host_graph" (1) / "dual_stream" (1)   eab_enhance_host_batches: replay each slot's step from a CUDA graph / alternate
 *                       batches on two compute streams
 *   "stream_tcm" (1)    streaming: the TCM stack as one launch
 *   "stream_umma" (1)   streaming: the per-layer convs on the tcgen05 gather kernel (all streams in one GEMM row space); 0 = CUDA cores
 *   "stream_fuse" (1)   streaming (BatchNorm models): the residual sum of a U-Net module in the epilogue of its last inner deconv instead
 *                       of a launch of its own (9 per step); changes the state layout like the kernel-selection options
 *   "stream_pair" (1)   streaming: the two output parities of a transposed conv as ONE launch (25 of the 81 conv launches of a step)
 *   "lstm_pp" (0)       offline LSTM layers on lstm_pp.cu: 128 sequences per CTA as two interleaved sub-batches whose gate GEMMs run
 *                       under each other's cell phase (tcgen05.ld.16x256b keeps all four schedulers' SFUs busy).  Same results to
 *                       rounding; measured 5.3 ms against 4.1 ms for both layers at 64 x 6 s (tensor-pipe bound), hence off
 *   "stream_lstm" (1)   streaming: the LSTM step as a tensor-core gate GEMM + LayerNorm / cell kernels (needs stream_umma)
 *   "norm_log" (0)      record the InstanceNorm statistics of the next eab_forward for eab_norm_stats (BN calibration)
 *   "raw_grid" (0), "dbg_launch" (-1), "lstm_exp" (0)   diagnostics (grid cap of conv_raw so that small inputs walk many tiles per
 *                       CTA; which tensor-core conv launch fills eab_debug_counters; LSTM ablation switches of debug builds)
 * Changing an option invalidates the CUDA graphs eab_enhance_host_batches has captured (they are re-captured on next use). */
EAB_API int     eab_set_option(eab_model* m, const char* name, int value);
/* diagnostics: cycle counters of CTA 0 of the tcgen05 conv launch selected with option "dbg_launch" */
EAB_API int     eab_debug_counters(eab_model* m, unsigned long long* out16);

/* Per-launch CUDA-event timing of the calling thread's launches, aggregated per kernel family; the summary is a
 * JSON array [{"kernel","launches","ms","flops","bytes","moved_bytes"}] - ALGORITHMIC flops/bytes in SURVEY.md 8(d)'s sense
 * (every fused layer reads its fp32 inputs once and writes its output once) and the bytes the launches really request -, written to
 * `buf` (returns its length, or -1).  Reading the summary synchronises the device and clears the records. */
EAB_API int     eab_profile_enable(eab_model* m, int on);
EAB_API int64_t eab_profile_summary(eab_model* m, char* buf, int64_t capacity);

EAB_API const char* eab_last_error(void);
EAB_API const char* eab_build_info(void);   /* "sm_100a;<date>;<nvcc>" */

#ifdef __cplusplus
}
#endif
#endif /* EABNET_B200_H_ */
