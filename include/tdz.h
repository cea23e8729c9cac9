/* tdz.h - C ABI of libtdz.so, the B200 (sm_100a) implementation of TargetDiarization's
 * target-speaker separation + scoring stage.
 *
 * The reference has no FFI: its seam for this path is two Python callables and one method
 * (SURVEY.md section 8b).  Each entry point below states which reference call it stands behind; the
 * Python host code in targetdiarization_b200/ mirrors those callables and reaches this library
 * through ctypes (INTEGRATION.md shows the binding).
 *
 *   AudioProcessor.separater(tensor[1,T]) -> [1,2,T]          AudioProcessor.py:271-273,943
 *       -> tdz_separate()                                      (look2hear/models/mossformer2.py:563-589)
 *   look2hear.utils.wav_chunk_inference (overlap-add)          look2hear/utils/separator.py:72-132
 *       -> tdz_gather_segments() + tdz_separate() + tdz_stitch_ola()
 *   AudioProcessor.separate_speaker chunk concat               AudioProcessor.py:920-948
 *       -> tdz_stitch_concat()
 *   TargetASR.embedding['eres2netv2_large'](wav, output_emb)   TargetASR.py:102-103,155-163
 *       -> tdz_fbank() + tdz_embed()                           (modelscope ERes2NetV2 pipeline)
 *   TargetASR.cosine_similarity                                TargetASR.py:144-152
 *       -> tdz_cosine_scores()
 *   ConvTDFNet.stft / .istft (MDX-Net front / back end)        AudioProcessor.py:82-120
 *       -> tdz_stft() / tdz_istft()
 *   AudioProcessor.restorer(tensor[1,nch,T]) (Apollo)          AudioProcessor.py:279,970
 *       -> tdz_apollo_restore()                                (look2hear/models/apollo.py:278-297)
 *
 * Conventions: all pointers named *_dev are device pointers owned by the caller (PyTorch allocates);
 * the library never allocates on the hot path, the caller passes a workspace sized by the matching
 * *_workspace_bytes().  Every call returns 0 on success or a non-zero code, with text in
 * tdz_last_error(); there is no CPU fallback.  `stream` is a cudaStream_t passed as void*.
 * A handle serialises its own calls; different handles may be used from different threads.  A workspace
 * is scratch memory of one call at a time: calls that share a workspace must be ordered by the caller (same
 * stream, or an event between streams) - the Python host objects do that (targetdiarization_b200/_lib.py::CallGuard).
 * Small calls (at most 32 768 padded frames per tdz_separate, 4 096 fbank frames per tdz_embed) launch their kernels
 * with the programmatic-stream-serialization attribute; every kernel of the library orders itself behind its
 * producer with griddepcontrol.wait, so this needs nothing from the caller (environment: TDZ_NO_PDL=1 turns it off).
 */
#ifndef TDZ_H_
#define TDZ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tdz_ctx tdz_ctx;

#define TDZ_NUM_LAYERS 24

/* Packed weights of one FLASH_ShareA_FFConvM + GatedFSMNBlockDilated pair
 * (look2hear/models/mossformer_block.py:143-220, 391-425; fsmn.py:76-144).  Norm gains that commute
 * with the following matmul are folded into it by the packer (targetdiarization_b200/weights.py). */
typedef struct tdz_layer_weights {
  const void* w_in;        /* bf16 [2176][512]: to_hidden (2048 rows, ScaleNorm g folded) | to_qk (128 rows) */
  const float* b_in;       /* [2176] */
  const float* dw_in;      /* [2176][20] per channel: the 17 depthwise taps of its ConvModule with +1 added to the
                              centre tap (x + conv(x), conv_module.py:219), then b_in / 2, then two zeros */
  const float* os_gamma;   /* [4][128] OffsetScale */
  const float* os_beta;    /* [4][128] */
  const void* w_out;       /* bf16 [512][1024], ScaleNorm g folded */
  const float* b_out;      /* [512] */
  const float* dw_out;     /* [512][20], same layout as dw_in */
  const float* w_c1;       /* fp32 [256][512] fsmn.conv1 (tf32 operand) */
  const float* b_c1;       /* [256] */
  const float* prelu_c1;   /* [1] */
  const float* ln1_g;      /* [256] norm1 */
  const float* ln1_b;      /* [256] */
  const void* w_uv;        /* bf16 [512][256]: to_u | to_v with their LayerNorm affine folded */
  const float* b_uv;       /* [512] */
  const float* dw_uv;      /* [512][20], same layout as dw_in */
  const void* w_lin;       /* bf16 [256][256] fsmn.linear */
  const float* b_lin;      /* [256] */
  const void* w_proj;      /* bf16 [256][256] fsmn.project (no bias) */
  const float* dd_w1;      /* [256][39] DilatedDenseNet conv1 */
  const float* in1_g;      /* [256] InstanceNorm affine */
  const float* in1_b;
  const float* dd_prelu1;  /* [256] */
  const float* dd_w2;      /* [256][2][39] conv2 (dilation 2) */
  const float* in2_g;
  const float* in2_b;
  const float* dd_prelu2;
  const float* w_c2;       /* fp32 [512][256] fsmn.conv2 with norm2 gain folded (tf32 operand) */
  const float* b_c2;       /* [512] with norm2 bias folded */
} tdz_layer_weights;

/* Whole MossFormer2 (look2hear/models/mossformer2.py:525-589), default architecture:
 * 512 channels, 24 blocks, kernel 16 / stride 8, 2 speakers. */
typedef struct tdz_mossformer2_weights {
  const float* enc_w;        /* [512][16] enc.conv1d */
  const float* w_enc1x1;     /* fp32 [512][512] mask_net.conv1d_encoder with GroupNorm gain folded */
  const float* enc1x1_colsum;/* [512] row sums of the folded matrix (GroupNorm mean term) */
  const float* enc1x1_bias;  /* [512] W @ beta (GroupNorm bias term) */
  const float* pos_inv_freq; /* [256] */
  const float* pos_scale;    /* [1] */
  const float* rot_freqs;    /* [16] rotary_pos_emb.freqs */
  tdz_layer_weights layers[TDZ_NUM_LAYERS];
  const float* fln_g;        /* [512] mdl.intra_mdl.norm (LayerNorm eps 1e-6) */
  const float* fln_b;
  const float* fgn_g;        /* [512] mdl.intra_norm (GroupNorm eps 1e-8) */
  const float* fgn_b;
  const float* mask_prelu;   /* [1] */
  const float* w_out1;       /* fp32 [1024][512] conv1d_out */
  const float* b_out1;       /* [1024] */
  const float* w_tg;         /* fp32 [1024][512]: output.0 (tanh) rows 0..511 | output_gate.0 rows 512..1023 */
  const float* b_tg;         /* [1024] */
  const float* w_dec1;       /* fp32 [512][512] conv1_decoder */
  const float* dec_w;        /* [512][16] dec (ConvTranspose1d) */
  const float* dec_wt;       /* fp32 [16][512] the same taps transposed, rounded to tf32 (N operand of the decoder GEMM) */
} tdz_mossformer2_weights;

/* ---- handle ---------------------------------------------------------------------------------- */
/* A handle is bound to `device`: every compute entry point makes that device current for the duration of the
 * call and restores the caller's current device, so handles on several GPUs may be driven from one thread. The
 * pointers and the stream passed to a call must belong to the handle's device. */
int tdz_create(int device, tdz_ctx** out);
void tdz_destroy(tdz_ctx* ctx);
const char* tdz_last_error(tdz_ctx* ctx);
int tdz_num_sms(tdz_ctx* ctx);
const char* tdz_version(void);

/* ---- separator ------------------------------------------------------------------------------- */
/* Stores the pointer table (the tensors stay owned by the caller and must outlive the handle's use). */
int tdz_set_mossformer2_weights(tdz_ctx* ctx, const tdz_mossformer2_weights* w);

/* Frames produced by the encoder for T samples: floor((T-16)/8)+1 (mossformer2.py:178-185). */
int64_t tdz_num_frames(int64_t T);
/* Frames per sample in the padded token space: S rounded up to a multiple of 256 (group size). */
int64_t tdz_padded_frames(int64_t T);
size_t tdz_separate_workspace_bytes(int64_t B, int64_t T);

/* MossFormer2.forward on B chunks of T samples each: mix_dev fp32 [B][T] -> out_dev fp32 [B][2][T].
 * Replaces `self.separater(audio_data_tensor)` (AudioProcessor.py:943). */
int tdz_separate(tdz_ctx* ctx, const float* mix_dev, int64_t B, int64_t T, float* out_dev, void* workspace_dev,
                 size_t workspace_bytes, void* stream);
/* Same forward with a strided output: stream `spk` of chunk `b` is written to
 * out_dev + b*out_chunk_stride + spk*out_spk_stride (strides in floats).  With out_chunk_stride = T and
 * out_spk_stride = the length of the stitched stream, adjacent equal-length windows land directly in the
 * concatenated [2][L] layout of AudioProcessor.separate_speaker (AudioProcessor.py:947-948): no stitch copy. */
int tdz_separate_strided(tdz_ctx* ctx, const float* mix_dev, int64_t B, int64_t T, float* out_dev,
                         int64_t out_chunk_stride, int64_t out_spk_stride, void* workspace_dev, size_t workspace_bytes,
                         void* stream);

/* Test hook: run only the first `num_layers` layer pairs, and of the launch sequence only the steps whose
 * index lies in [step_lo, step_hi] (step numbering: enum Step in csrc/tdz_api.cu; 0..20).  The tests write
 * oracle intermediates into the workspace, run one step and compare its outputs.  Offsets (bytes) of the
 * named intermediates inside the workspace are returned by tdz_separate_layout(). */
int tdz_separate_debug(tdz_ctx* ctx, const float* mix_dev, int64_t B, int64_t T, float* out_dev, void* workspace_dev,
                       size_t workspace_bytes, void* stream, int num_layers, int step_lo, int step_hi);

typedef struct tdz_sep_layout {
  size_t enc, x0, x, xbf, ss, vu, qk4, lq_lo, qkf, P, o, o_ss, c, nhat, xuv, xubf, f1, p, y1, y2, g, lnb, ab, mb, gated, sep,
      kv_part, kv, gn_stats, in_stats, in_ss, samp, rot, hrs, pos, total;
  int64_t S, Sp, Mtot;
  int32_t kv_nsplit, kv_kb_per_split;
} tdz_sep_layout;
int tdz_separate_layout(int64_t B, int64_t T, int num_sms, tdz_sep_layout* out);

/* ---- chunk stitching ------------------------------------------------------------------------- */
/* wav_chunk_inference segmenting (separator.py:95-112): seg_dev[i][:] = padded_mix[i*hop : i*hop+session],
 * zero where outside; padded_mix = zeros(session-hop) | mix | zeros(session-hop).  Segments
 * [seg_begin, seg_begin+n_seg). */
int tdz_gather_segments(tdz_ctx* ctx, const float* mix_dev, int64_t L, int64_t session, int64_t hop,
                        int64_t seg_begin, int64_t n_seg, float* seg_dev, void* stream);
/* Same for a rank that holds only samples [mix_origin, mix_origin + mix_len) of the L-sample mixture (its output
 * span plus the halo its segments reach into): mix_dev[0] is sample mix_origin.  Samples
 * outside [0, L) are zeros as above; asking for a sample inside [0, L) but outside the resident range is an error
 * caught on the host (the segment range is checked against the resident range before the launch). */
int tdz_gather_segments_span(tdz_ctx* ctx, const float* mix_dev, int64_t mix_origin, int64_t mix_len, int64_t L,
                             int64_t session, int64_t hop, int64_t seg_begin, int64_t n_seg, float* seg_dev,
                             void* stream);
/* Overlap-add (separator.py:126-130): out[trk][n] = (1/ratio) * sum_i est[i][trk][n + (session-hop) - i*hop]
 * over the segments covering n, ascending i (fixed order -> bit-reproducible across shardings).
 * est_dev fp32 [n_seg][2][session] holds segments [seg_begin, seg_begin+n_seg); out_dev fp32 [2][n_out]
 * receives samples [out_begin, out_begin+n_out) of the stitched streams. */
int tdz_stitch_ola(tdz_ctx* ctx, const float* est_dev, int64_t session, int64_t hop, int64_t seg_begin,
                   int64_t n_seg, int64_t L, int64_t out_begin, int64_t n_out, float ratio, float* out_dev,
                   void* stream);
/* separate_speaker concat (AudioProcessor.py:947-948): copies est_dev [2][len] of one chunk into
 * out_dev [2][L] at sample offset `start`. */
int tdz_stitch_concat(tdz_ctx* ctx, const float* est_dev, int64_t len, int64_t start, int64_t L, float* out_dev,
                      void* stream);

/* ---- loudness -------------------------------------------------------------------------------- */
/* The heavy part of AudioProcessor.meter_loudness (AudioProcessor.py:1123-1127; pyloudnorm integrated loudness,
 * BS.1770): K-weighting of n_streams signals x_dev fp32 [n_streams][L] in fp64 (coef = for each of the two biquads
 * b0 b1 b2 a0 a1 a2, a0 == 1) and the mean square of every 400 ms block [lo[j], hi[j]) -> z_dev fp64
 * [n_streams][nblk] (= sum * inv_len).  ysq_dev is fp64 scratch [n_streams][L].  The two gates over the block
 * values (a few thousand numbers) are applied by the caller. */
int tdz_loudness_blocks(tdz_ctx* ctx, const float* x_dev, int64_t n_streams, int64_t L, const double* coef,
                        const int64_t* lo_dev, const int64_t* hi_dev, int64_t nblk, double inv_len, double* ysq_dev,
                        double* z_dev, void* stream);

/* ---- speaker scoring ------------------------------------------------------------------------- */
/* Kaldi fbank (80 mel bins, 25 ms / 10 ms, povey window, pre-emphasis 0.97, DC removal, power, log,
 * snip_edges) + per-utterance mean normalisation, as the modelscope ERes2NetV2 pipeline computes it
 * (torchaudio/compliance/kaldi.py:514-646).  wav_dev fp32 [N][T] -> feat_dev fp32 [N][frames][80],
 * frames = 1 + (T-400)/160. */
int64_t tdz_fbank_frames(int64_t T);
/* Constant tables (povey window [400], FFT twiddles [256][2], mel filters [80][257], first/last non-zero
 * bin per filter [80]) computed by targetdiarization_b200/fbank.py; device pointers, caller-owned. */
int tdz_set_fbank_tables(tdz_ctx* ctx, const float* window_dev, const float* twiddle_dev, const float* mel_dev,
                         const int32_t* mel_lo_dev, const int32_t* mel_hi_dev);
int tdz_fbank(tdz_ctx* ctx, const float* wav_dev, int64_t N, int64_t T, float* feat_dev, void* stream);

/* ERes2NetV2-Large (modelscope `ERes2NetV2`, m_channels 64, blocks [3,4,6,3], baseWidth 24, scale 4,
 * expansion 4, TSTP pooling, 192-d embedding).  Every convolution has its eval-mode
 * BatchNorm folded in by the packer (targetdiarization_b200/weights.py) and is stored as a GEMM operand:
 * w = bf16 [round_up(N, BN)][round_up(K, 64)], zero padded, K ordered (kh, kw, cin) for 3x3 kernels;
 * b = fp32 [round_up(N, BN)];  BN = 32 / 64 / 128 / 256 = the smallest of these >= N. */
#define TDZ_SV_NUM_BLOCKS 16
#define TDZ_SV_EMBED_DIM 192
typedef struct tdz_conv {
  const void* w;
  const float* b;
} tdz_conv;
typedef struct tdz_eres_block {
  tdz_conv conv1;      /* 1x1, in_planes -> 4*width (stride on this conv) */
  tdz_conv convs[4];   /* 3x3, width -> width */
  tdz_conv aff_a[3];   /* AFF blocks only (layers 3-4): 1x1, 2*width -> width/4 */
  tdz_conv aff_b[3];   /*                               1x1, width/4 -> width */
  tdz_conv conv3;      /* 1x1, 4*width -> 4*planes */
  tdz_conv shortcut;   /* 1x1 (stride), in_planes -> 4*planes; w == NULL for identity shortcuts */
} tdz_eres_block;
typedef struct tdz_eres2netv2_weights {
  const float* stem_w; /* fp32 [64][9] conv1 + bn1 folded */
  const float* stem_b; /* fp32 [64] */
  tdz_eres_block blocks[TDZ_SV_NUM_BLOCKS];
  tdz_conv layer3_ds;  /* 3x3 stride 2, 1024 -> 2048, no BN */
  tdz_conv fuse_a;     /* fuse34 AFF: 4096 -> 512 */
  tdz_conv fuse_b;     /*             512 -> 2048 */
  tdz_conv seg1;       /* Linear 40960 -> 192 */
} tdz_eres2netv2_weights;
int tdz_set_eres2netv2_weights(tdz_ctx* ctx, const tdz_eres2netv2_weights* w);
size_t tdz_embed_workspace_bytes(int64_t N, int64_t frames);
/* feat_dev fp32 [N][frames][80] (output of tdz_fbank) -> emb_dev fp32 [N][192].  Replaces the model call
 * inside `self.embedding['eres2netv2_large'](wav, output_emb=True)` (TargetASR.py:161). */
int tdz_embed(tdz_ctx* ctx, const float* feat_dev, int64_t N, int64_t frames, float* emb_dev, void* workspace_dev,
              size_t workspace_bytes, void* stream);

/* Test hook: stops after the stem (stop_block -1), after residual block stop_block (0..15) or after the fuse34
 * map (16) and copies that NHWC feature map [N*H*W][C] into out_dev (sized by the caller): bf16 for the stem and
 * the blocks (feature maps are stored in bf16), fp32 for fuse34. */
int tdz_embed_debug(tdz_ctx* ctx, const float* feat_dev, int64_t N, int64_t frames, float* out_dev,
                    void* workspace_dev, size_t workspace_bytes, void* stream, int stop_block);

/* cosine similarity of N embeddings against one target, TargetASR.cosine_similarity semantics
 * (all-zero vector -> 1.0; result clamped to [0,1]).  emb_dev [N][dim], target_dev [dim] -> scores_dev [N]. */
int tdz_cosine_scores(tdz_ctx* ctx, const float* emb_dev, const float* target_dev, int64_t N, int64_t dim,
                      float* scores_dev, void* stream);

/* ---- STFT / inverse STFT -------------------------------------------------------------
 * torch.stft / torch.istft with the arguments the reference passes (center = True, reflect padding, periodic Hann
 * window, one-sided, un-normalised): ConvTDFNet.stft / .istft (AudioProcessor.py:82-120, n_fft 6144) and the Apollo
 * restorer (look2hear/models/apollo.py:261-262, 294-295, n_fft 882).  n_fft must factor into 2, 3, 5, 7.
 * The spectrogram element (row r, bin k, frame t, re/im c) lives at spec_dev[r*s_row + k*s_bin + t*s_frame + c*s_reim],
 * so the MDX layout [B, 4 = channel*2 + re/im, dim_f, dim_t] is (2*dim_f*dim_t, dim_t, 1, dim_f*dim_t) with r = 2*b +
 * channel, and frame-major complex [rows, T, bins] is (T*bins*2, 2, bins*2, 1).  Bins [0, n_keep) are written / read;
 * tdz_istft takes the others as zero (the freq_pad of ConvTDFNet.istft). */
typedef struct tdz_stft_plan {
  int32_t n_fft, hop;
  const float* window_dev;   /* [n_fft] */
  const float* twiddle_dev;  /* [n_fft][2]: cos(2 pi k / n_fft), -sin(2 pi k / n_fft), rounded from double */
} tdz_stft_plan;
int64_t tdz_stft_frames(int64_t L, int64_t hop);   /* 1 + L / hop */
/* x_dev fp32 [rows][L] -> spectrogram of tdz_stft_frames(L, hop) frames.  Needs n_fft / 2 < L (reflect padding). */
int tdz_stft(tdz_ctx* ctx, const tdz_stft_plan* plan, const float* x_dev, int64_t rows, int64_t L, int64_t n_keep,
             float* spec_dev, int64_t s_row, int64_t s_bin, int64_t s_frame, int64_t s_reim, void* stream);
/* spectrogram of T frames -> out_dev fp32 [rows][out_len] (torch.istft(..., length = out_len); without `length` the
 * reference gets out_len = hop * (T - 1)).  frames_ws_dev: fp32 scratch of rows * T * n_fft values. */
int tdz_istft(tdz_ctx* ctx, const tdz_stft_plan* plan, const float* spec_dev, int64_t rows, int64_t T, int64_t n_keep,
              int64_t s_row, int64_t s_bin, int64_t s_frame, int64_t s_reim, float* frames_ws_dev, float* out_dev,
              int64_t out_len, void* stream);

/* ---- Apollo restorer (look2hear/models/apollo.py; AudioProcessor.restore_audio, AudioProcessor.py:959-980) ------
 * The architecture AudioProcessor.init_restorer_model builds (AudioProcessor.py:279): Apollo(sr 44100, win 20 ms ->
 * n_fft 882 / hop 441, feature_dim 256, layer 6), 80 bands (79 x 5 bins + 47).  GEMM operands are bf16 [N][K]
 * row-major with the RMSNorm gain in front of a conv folded into its columns (targetdiarization_b200/weights.py). */
#define TDZ_AP_LAYERS 6
#define TDZ_AP_NBAND 80
#define TDZ_AP_BINS 442
typedef struct tdz_apollo_icb {
  const float* dw;     /* [7][256] depthwise taps, tap-major */
  const float* dw_b;   /* [256] */
  const void* w1;      /* bf16 [1024][256], RMSNorm gain folded */
  const float* b1;     /* [1024] */
  const void* w2;      /* bf16 [256][1024] */
  const float* b2;     /* [256] */
} tdz_apollo_icb;
typedef struct tdz_apollo_layer {
  const void* w_qkv;   /* bf16 [768][256]  band_net.weight, input_norm gain folded; row = head*96 + (q|k|v)*32 + d */
  const void* w_out;   /* bf16 [256][256]  band_net.output */
  const void* w_mlp1;  /* bf16 [2048][256] band_net.MLP.1 (rows 0..1023 gate, 1024..2047 z), MLP.0 gain folded */
  const void* w_mlp2;  /* bf16 [256][1024] band_net.MLP_output */
  tdz_apollo_icb icb[3];
} tdz_apollo_layer;
typedef struct tdz_apollo_weights {
  const float* bn_g;     /* [964]      BN[i].0 gains, bands concatenated (2*BW+1 each) */
  const float* bn_w;     /* [964][256] BN[i].1 weights transposed: (band, input feature)-major, output contiguous */
  const float* bn_b;     /* [80][256] */
  const float* rot_cos;  /* [100][32]  band_net.cos_freq (identical in every layer) */
  const float* rot_sin;  /* [100][32] */
  tdz_apollo_layer layers[TDZ_AP_LAYERS];
  const float* out_g;    /* [80][256]  output[i].0 gains */
  const float* out_wv;   /* [256][884] output[i].1 rows that GLU keeps (real BW | imag BW per band), transposed */
  const float* out_wg;   /* [256][884] the matching gate rows, transposed */
  const float* out_bv;   /* [884] */
  const float* out_bg;   /* [884] */
  tdz_stft_plan plan;    /* n_fft 882, hop 441 */
} tdz_apollo_weights;
int tdz_set_apollo_weights(tdz_ctx* ctx, const tdz_apollo_weights* w);
/* Workspace of a single pass over the whole input.  Any size from tdz_apollo_min_workspace_bytes() up is accepted:
 * with less than the single-pass size the network runs over frame chunks with a 54-frame halo (the receptive field of
 * the 54 depthwise k7 convolutions) - same bits, bounded memory (the single pass needs 493 KB per 10 ms frame). */
size_t tdz_apollo_workspace_bytes(int64_t rows, int64_t nsample);
size_t tdz_apollo_min_workspace_bytes(int64_t rows, int64_t nsample);
/* self.restorer(tensor[B, nch, nsample]) (AudioProcessor.py:970; Apollo.forward, apollo.py:278-297):
 * wav_dev fp32 [rows = B * nch][nsample] -> out_dev fp32 [rows][nsample]. */
int tdz_apollo_restore(tdz_ctx* ctx, const float* wav_dev, int64_t rows, int64_t nsample, float* out_dev,
                       void* workspace_dev, size_t workspace_bytes, void* stream);
/* Test hook: runs up to a tap and copies it to out_dev (fp32): 0 = spectrogram [rows*T][442][2]; 1 = band-split
 * features [tokens][256]; 2 = attention output of layer 0 (bf16 widened) [tokens][256]; 3 = layer 0 after the
 * Roformer [tokens][256]; 4 + l = output of layer l; 10 = estimated spectrogram [rows*T][442][2]. */
int tdz_apollo_debug(tdz_ctx* ctx, const float* wav_dev, int64_t rows, int64_t nsample, float* out_dev,
                     void* workspace_dev, size_t workspace_bytes, void* stream, int tap);

#ifdef __cplusplus
}
#endif
#endif /* TDZ_H_ */
