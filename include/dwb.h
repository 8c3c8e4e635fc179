/* distil-whisper-b200: C ABI of the sm_100a kernel library (libdwb.so).
 *
 * The reference (huggingface/distil-whisper @ cc96130) has no native code and no FFI: its hot path is
 * training/run_distillation.py:1465-1495 (train_step) calling Hugging Face Transformers' Whisper modules.  Each
 * entry point below replaces one piece of arithmetic that path reaches; the file:line it replaces is cited per
 * function ("ref:" = the reference repo, "HF:" = transformers 5.5.0, the version the oracle is pinned to).
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (DWB_ERR_*); dwb_last_error() gives the text (thread local).
 *   - all pointers are DEVICE pointers unless named *_host; buffers are caller-owned, never freed or allocated here.
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on that stream and re-entrant.
 *   - "bf16" buffers are __nv_bfloat16; matrices are row-major with an explicit row pitch `ld*` in ELEMENTS.
 *   - TMA-backed operands (GEMM A/B/C) need a 16-byte aligned base and a row pitch that is a multiple of 16 bytes.
 *   - no hidden global state except immutable tables created by *_plan_create (the tensor-core log-mel plan also owns two
 *     internal streams + events for its chunk pipeline: one dwb_logmel_tc call at a time per plan), the launch counter and the
 *     optimiser-tail grid cap.
 *   - host-side contract of the Python layer built on this ABI (distil_whisper_b200/): parameter gradients are written by these kernels
 *     straight into caller-owned flat fp32 buffers (TMA reduce-add / atomics), never through autograd hooks -- a data-parallel user
 *     all-reduces that buffer once per optimiser step (optim.FusedAdamW.all_reduce_gradients) instead of wrapping the model in DDP.
 */
#ifndef DWB_H_
#define DWB_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DWB_OK 0
#define DWB_ERR_INVALID (-1)
#define DWB_ERR_CUDA (-2)
#define DWB_ERR_UNSUPPORTED (-3)

const char* dwb_last_error(void);
int dwb_abi_version(void);
/* Kernel launches issued by the library since the last reset (host-side count; bench.py's `gpu_launches`). */
int64_t dwb_launch_count(int reset);
/* Direction in which the NEXT dwb_gemm_bf16 / dwb_add_layernorm / dwb_attention_fwd_tc launches walk the rows of their operands
 * (0 = first to last, the default; 1 = last to first).  Alternating it along a chain of kernels that stream tensors larger than L2 makes
 * every consumer start on the rows its producer wrote last.  Results are unaffected. */
int dwb_set_row_walk(int reverse);
/* 0 iff the current CUDA device is compute capability 10.x (B200).  There is no CPU fallback. */
int dwb_check_device(void);

/* ---- dense contractions (tcgen05 + TMA + TMEM) -------------------------------------------------------------
 * C[M,N] = act(alpha * A . B^T + bias)            bf16 inputs, fp32 accumulate, C bf16 (c_f32=0) or fp32 (c_f32=1)
 *   a_mn_major = 0: A is [M,K] (pitch lda)   1: A is stored [K,M] (pitch lda)
 *   b_mn_major = 0: B is [N,K] (pitch ldb)   1: B is stored [K,N] (pitch ldb)
 *   act: 0 none, 1 exact-erf GELU.  accumulate=1: C += result (fp32 C only; TMA reduce-add).
 *   impl: 0 = tcgen05 kernel (product path), 1 = plain SIMT kernel (cross-check only).
 * Replaces nn.Linear forward/backward at HF:models/whisper/modeling_whisper.py:310-355 (q/k/v/out_proj), :404-407
 * and :497-500 (fc1 / gelu / fc2), :1081 (proj_out), and the two Conv1d of :619-620 after im2col. */
int dwb_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major, void* C, int64_t ldc,
                  int c_f32, int M, int N, int K, const float* bias, int act, float alpha, int accumulate, int impl, void* stream);

/* ---- attention (head_dim 64) ---------------------------------------------------------------------------------
 * O = softmax(scale * Q K^T [+ causal mask]) V per (batch, head); LSE[b,h,i] = log sum_j exp(scale * q_i.k_j).
 * Q rows are [B*Sq, ldq] with head h at columns [64h, 64h+64); likewise K, V ([B*Sk, .]) and O.
 * Replaces F.scaled_dot_product_attention via HF:integrations/sdpa_attention.py:40-104 (called from
 * HF:models/whisper/modeling_whisper.py:342-352 with scaling=1.0 and q pre-scaled at :310; pass scale = 64^-0.5). */
int dwb_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo,
                      float* lse, int B, int H, int Sq, int Sk, int head_dim, int causal, float scale, void* stream);
/* tcgen05 / TMEM forward for the non-causal encoder self-attention (same contract, causal must be 0). */
int dwb_attention_fwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o, int64_t ldo,
                         float* lse, int B, int H, int Sq, int Sk, int head_dim, int causal, float scale, void* stream);
/* Backward.  delta_ws: fp32 [B*H*Sq] scratch; dq_acc: fp32 [B*Sq, H*64] (zeroed here, accumulated with atomics);
 * dk/dv: bf16 with pitches lddk/lddv.  Autograd of the same sdpa call (ref:training/run_distillation.py:1609). */
int dwb_attention_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* o, int64_t ldo,
                      const void* dout, int64_t lddo, const float* lse, float* delta_ws, float* dq_acc, void* dk, int64_t lddk,
                      void* dv, int64_t lddv, int B, int H, int Sq, int Sk, int head_dim, int causal, float scale, void* stream);

/* Same contract on tcgen05 / TMEM (all five contractions of the flash-attention backward on the tensor core). */
int dwb_attention_bwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* o, int64_t ldo,
                         const void* dout, int64_t lddo, const float* lse, float* delta_ws, float* dq_acc, void* dk, int64_t lddk,
                         void* dv, int64_t lddv, int B, int H, int Sq, int Sk, int head_dim, int causal, float scale, void* stream);

/* ---- residual add + LayerNorm ---------------------------------------------------------------------------------
 * x_new[r,:] = x_in[r % x_rows_mod (0: r), :] + y[r,:] (y bf16, nullable); ln = LayerNorm(x_new; gamma, beta, eps).
 * x_out (fp32), ln_out (bf16), mean/rstd (fp32 [rows]) are each optional.
 * HF:models/whisper/modeling_whisper.py:392-409, :469-503 (residual + next pre-LN), :623-625 + :643, :791. */
int dwb_add_layernorm(const float* x_in, int x_rows_mod, const void* y_bf16, const float* gamma, const float* beta, float* x_out,
                      void* ln_out_bf16, float* mean_out, float* rstd_out, int rows, int d, float eps, void* stream);
/* dx = dres + dLN(dy); optionally also as bf16; dgamma/dbeta are ACCUMULATED (atomicAdd). */
int dwb_layernorm_bwd(const void* dy_bf16, const float* x, const float* mean, const float* rstd, const float* gamma,
                      const float* dres, float* dx, void* dx_bf16, float* dgamma, float* dbeta, int rows, int d, void* stream);

/* ---- casts / layout -------------------------------------------------------------------------------------------*/
int dwb_cast_f32_to_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int rows, int cols, float scale, void* stream);
int dwb_cast_bf16_to_f32(const void* src, int64_t lds, float* dst, int64_t ldd, int rows, int cols, void* stream);
/* x *= *scale_dev (device scalar; no-op pass when it is 1): applies the upstream gradient of `loss.backward()` /
 * accelerator.backward(loss) (ref:training/run_distillation.py:1609, which divides by gradient_accumulation_steps) to the
 * bf16 d loss / d logits produced by dwb_kd_loss.  n % 8 == 0, contiguous. */
int dwb_scale_bf16_dev(void* x_bf16, int64_t n, const float* scale_dev, void* stream);
/* Conv1d weight [O,C,3] fp32 -> bf16 [O,3C] with column k*C+c (conv2), and the inverse for its gradient. */
int dwb_conv_weight_to_kc_bf16(const float* w, void* out_bf16, int O, int C, void* stream);
int dwb_conv_wgrad_kc_to_ck(const float* g, float* dw, int O, int C, int accumulate, void* stream);
/* im2col for HF:models/whisper/modeling_whisper.py:619 (conv1: mel [B,C,L] fp32 -> [B*L, ld] bf16, col c*3+k) and
 * :620 (conv2 on channels-last x [B,L,d] bf16 -> [B*L/2, 3d], col k*d+c). */
int dwb_im2col_conv1(const float* mel, void* out_bf16, int B, int C, int L, int ld, void* stream);
int dwb_im2col_conv2(const void* x_bf16, void* out_bf16, int B, int L, int d, void* stream);
/* Backward of :620/:619: col2im of the conv2 input gradient g [B*L/2, 3d] onto [B, L, d], times gelu'(conv1 pre-activation). */
int dwb_col2im_conv2_gelu_bwd(const void* g_bf16, const void* pre1_bf16, void* out_bf16, int B, int L, int d, void* stream);

/* ---- decoder embeddings: HF:models/whisper/modeling_whisper.py:738 (embed_tokens, padding_idx) + :755 (positions) */
int dwb_embed_fwd(const int64_t* ids, const void* E, const void* P, int table_is_f32, float* x, int B, int T, int d, int vocab,
                  void* stream);
int dwb_embed_bwd(const int64_t* ids, const float* dx, float* dE, float* dP, int B, int T, int d, int vocab, int padding_idx,
                  void* stream);

/* ---- single-token decoder step (greedy generation with a KV cache) ---------------------------------------------------
 * The eval loop's `student_model.generate(batch["input_features"], **gen_kwargs)` (ref:training/run_distillation.py:1524-1528)
 * and the pseudo-labelling loop (ref:training/run_pseudo_labelling.py:861-927) decode one token per step against cached
 * keys / values (HF:models/whisper/modeling_whisper.py:315-340).  Every position-dependent quantity is read from device
 * memory (`pos_dev`: the index of the token being consumed), so one captured CUDA graph serves every step.
 *   seq [B, seq_ld] int64: prompt tokens then generated tokens.
 * dwb_embed_decode:      x[b,:] = E[seq[b,pos]] + P[pos]                      (HF :738, :755 with past length pos)
 * dwb_attention_decode:  one query row per (b, h) over cache rows [0, len): k_new/v_new non-null -> appended at row pos,
 *                        len = pos + 1 (self-attention); null -> len = fixed_len (cross-attention over the encoder states).
 *                        Cache: K and V rows of pitch ld_cache elements, `cache_rows` rows per batch entry.
 * dwb_greedy_pick:       next = argmax(logits + bias_all + [pos+1 == begin_pos] bias_begin)  (HF:generation/logits_process.py
 *                        SuppressTokensLogitsProcessor / SuppressTokensAtBeginLogitsProcessor as 0 / -inf biases), prompt
 *                        positions are kept, finished rows get `pad`, `finished` is updated on `eos` (HF:generation/utils.py _sample).
 * dwb_decode_advance:    pos += 1; done_at = sequence length at which every row had finished (0 until then). */
/* dwb_gemm_skinny_bf16: C[M,N] = act(X[M,K] . W[N,K]^T + bias) for the decode step's batch-sized M (16, 32, 48 or 64 rows; N % 8 == 0,
 * K % 256 == 0, 16 B aligned rows): weight-bandwidth bound, N / 8 CTAs of four warps (8 columns x 4 k-ranges), mma.sync.m16n8k16 (same
 * nn.Linear semantics as dwb_gemm_bf16). */
int dwb_gemm_skinny_bf16(const void* X, int64_t ldx, const void* W, int64_t ldw, void* C, int64_t ldc, int c_f32, int M, int N, int K,
                         const float* bias, int act, void* stream);
int dwb_embed_decode(const int64_t* seq, int seq_ld, const int* pos_dev, const void* E, const void* P, int table_is_f32, float* x, int B,
                     int d, int vocab, void* stream);
int dwb_attention_decode(const void* q, int64_t ldq, const void* k_new, const void* v_new, int64_t ld_new, void* k_cache, void* v_cache,
                         int64_t ld_cache, int cache_rows, void* o, int64_t ldo, int B, int H, int head_dim, int fixed_len,
                         const int* pos_dev, float scale, void* stream);
int dwb_greedy_pick(const float* logits, int64_t ld, int vocab, const float* bias_all, const float* bias_begin, int begin_pos, int64_t* seq,
                    int seq_ld, int prompt_len, int* finished, int64_t eos, int64_t pad, const int* pos_dev, int B, void* stream);
/* dwb_greedy_pick under Whisper's timestamp rules (return_timestamps=True, the reference's recommended pseudo-labelling mode,
 * ref:training/README.md:130,148): HF:generation/logits_process.py WhisperTimeStampLogitsProcessor applied after the suppress biases.
 * timestamp_begin = <|notimestamps|> + 1; max_initial_timestamp_index < 0: no limit. */
int dwb_greedy_pick_timestamps(const float* logits, int64_t ld, int vocab, const float* bias_all, const float* bias_begin, int begin_pos,
                               int64_t* seq, int seq_ld, int prompt_len, int* finished, int64_t eos, int64_t pad, const int* pos_dev, int B,
                               int timestamp_begin, int max_initial_timestamp_index, void* stream);
int dwb_decode_advance(int* pos_dev, const int* finished, int B, int* done_at, void* stream);

/* ---- label side of the data collator: ref:training/run_distillation.py:460-476 --------------------------------------
 * tokens [B, L1] int64 (padded), lengths [B] int32 -> decoder_input_ids [B, L1-1] = tokens[:, :-1]; labels [B, L1-1] =
 * tokens[:, 1:] with -100 on padding and on the prompt up to and including <|startoftranscript|>. */
int dwb_collate_labels(const int64_t* tokens, const int* lengths, int B, int L1, int64_t decoder_start_token_id,
                       int64_t* decoder_input_ids, int64_t* labels, void* stream);

/* ---- small reductions / activations ---------------------------------------------------------------------------*/
int dwb_colsum_bf16(const void* m_bf16, int64_t ld, float* out, int rows, int cols, int accumulate, void* stream); /* bias grads */
int dwb_gelu_bwd(const void* da, const void* h, void* dh, int64_t n, void* stream);
int dwb_gelu_fwd(const void* h, void* y, int64_t n, void* stream);

/* ---- KD loss head: ref:training/run_distillation.py:1453-1462 + :1484-1493, HF:...modeling_whisper.py:1085-1088 ----
 * metrics4 = {loss, ce, kl, n_valid} (device).  dlogits (bf16, pitch ldd, nullable) = d loss / d student_logits.
 * teacher_logits may be NULL (CE only).  workspace: dwb_kd_loss_workspace_bytes(rows). */
int64_t dwb_kd_loss_workspace_bytes(int rows);
int dwb_kd_loss(const float* student_logits, const float* teacher_logits, int64_t ld, const int64_t* labels, int rows, int vocab,
                float temperature, float ce_weight, float kl_weight, float* metrics4, void* dlogits_bf16, int64_t ldd,
                void* workspace, void* stream);

/* ---- optimiser tail: ref:training/run_distillation.py:1610-1614 (clip_grad_norm_, AdamW.step, zero_grad) -------*/
int dwb_grad_sumsq(const float* g, int64_t n, float* out_accum, void* stream);
/* Cap the grid of the two optimiser-tail kernels (0 = default: fill the machine).  Used when the tail is overlapped with the
 * next step's encoder forward on a side stream, so that it never takes an SM away from the persistent GEMM kernels. */
int dwb_set_tail_grid(int ctas);
int dwb_adamw_step(float* p, float* g, float* m, float* v, void* p_bf16, int64_t n, float lr, float beta1, float beta2, float eps,
                   float weight_decay, int step, const float* grad_sumsq, float max_grad_norm, float grad_scale, int zero_grad,
                   void* stream);

/* ---- the gradient all-reduce over symmetric (peer-mapped) memory: ref:training/run_distillation.py:1609 (DDP's implicit all-reduce) ---
 * In place, "two-shot": rank r sums slice r of the N copies (NVSwitch multimem.ld_reduce when multicast_ptr != NULL, else plain loads
 * through peer_ptrs[world]) and writes the sum into slice r of all N buffers.  The caller provides the cross-rank barriers before and
 * after (torch.distributed._symmetric_memory handle).  A few no-smem CTAs (max_ctas, default 32): co-resident with the GEMM kernels. */
int dwb_allreduce_symm(void* multicast_ptr, const void* const* peer_ptrs, int rank, int world, int64_t n, int max_ctas, void* stream);

/* ---- log-mel feature extractor: HF:models/whisper/feature_extraction_whisper.py:135-164 ------------------------
 * plan: mel filter bank [201, n_mels] fp32 on the HOST (HF:audio_utils.py:453-544) -> device tables.
 * wav [B, 480000] fp32 -> out [B, n_mels, 3000] fp32.  One 8-CTA cluster per utterance. */
int dwb_logmel_plan_create(const float* mel_filters_host, int n_freq, int n_mels, void** plan_out);
int dwb_logmel_plan_destroy(void* plan);
int dwb_logmel(void* plan, const float* wav, int B, int n_samples, float* out, void* stream);

/* Same contract on the tensor cores: the windowed 400-point DFT of every frame as a tcgen05 GEMM (fp16 hi/lo split of both
 * operands, three MMAs per k-step, fp32 accumulation in TMEM), the overlapping frames delivered by TMA through a tensor map with a
 * 160-sample frame stride, mel / log10 / per-utterance maximum in the epilogue.  workspace: dwb_logmel_tc_workspace_bytes(B,
 * n_samples) bytes, 256 B aligned, caller-owned (fp16 copies of one chunk of padded waveforms + per-utterance maxima). */
int dwb_logmel_tc_plan_create(const float* mel_filters_host, int n_freq, int n_mels, void** plan_out);
int dwb_logmel_tc_plan_destroy(void* plan);
int64_t dwb_logmel_tc_workspace_bytes(int B, int n_samples);
int dwb_logmel_tc(void* plan, const float* wav, int B, int n_samples, float* out, void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DWB_H_ */
