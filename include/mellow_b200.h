/* mellow_b200 -- C ABI of the B200-native Mellow inference path.
 *
 * The reference (soham97/mellow) has no FFI / plugin interface: its boundary is the Python class
 * `MellowWrapper` (mellow/wrapper.py:25-287).  This library replaces what sits *under* that class:
 *
 *   reference call site                                              replaced by
 *   ---------------------------------------------------------------  ------------------------------
 *   model.to(f"cuda:{device}")                 wrapper.py:87-88       mb_create + mb_bind_weights
 *   htsat.spectrogram_extractor/logmel/bn0     htsat.py:864-870       mb_frontend
 *   AudioEncoder.forward (x2 clips)            mellow.py:64-68,105-6  mb_encode
 *   DecoderModel.generate_prefix_inference     decoder.py:36-55       mb_prefix
 *   lm(inputs_embeds=prefix).logits[:, -1]     wrapper.py:217-218     mb_prefill
 *   the decode loop                            wrapper.py:216-249     mb_decode
 *   generate_prefix_inference + _generate_batch wrapper.py:285-286    mb_generate / mb_generate_host
 *
 * Conventions: every function returns 0 on success, non-zero on failure (message via mb_last_error).  Pointers are
 * DEVICE pointers unless the name ends in `_host`.  One handle per device; a handle is not thread-safe; no device
 * allocation happens after mb_create.  `stream` is a cudaStream_t passed as void* (NULL = default stream).
 * There is no CPU fallback: without a CUDA device mb_create fails.
 */
#ifndef MELLOW_B200_H
#define MELLOW_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define MB_POLICY_SPLIT 0 /* bf16 hi/lo operand split, 3 MMA passes, fp32 KV cache: greedy ids match the fp32 reference */
#define MB_POLICY_FAST 1  /* single bf16 operand plane, 1 MMA pass, bf16 KV cache: logits within bf16 tolerance */
#define MB_POLICY_SPLIT24 2 /* MB_POLICY_SPLIT with the KV cache rounded to 24 bits (bf16 upper half + one mantissa byte per value,
                             * relative error 2^-17 like the split GEMM operands): 25 % fewer bytes for decode attention */

int mb_version(void);

/* ---- weights: the library owns the arena layout; the host packer asks for it by entry name ---- */
int mb_weight_entry_count(void);
const char* mb_weight_entry_name(int i);
long long mb_weight_entry_offset(int i);
long long mb_weight_entry_bytes(int i);
long long mb_weights_size(void);

/* ---- lifetime ---- */
void* mb_create(int device, int max_batch, int max_new_tokens, int policy);
void mb_destroy(void* h);
const char* mb_last_error(void* h); /* h may be NULL for mb_create failures */
/* binds a packed device arena of mb_weights_size() bytes (caller keeps it alive; it is what gets NCCL-broadcast) */
int mb_bind_weights(void* h, const void* dev_arena, long long nbytes);
long long mb_workspace_bytes(void* h);
/* handle-scoped options (defaults in parentheses; none changes a result beyond fp32 summation order except
 * skip_finished, see mb_decode):
 *   "graph"          (1) replay the decode step as one CUDA graph; 0 = individual launches
 *   "decode_unfused" (0) 1 = use the generic per-layer decode path (the one batches > 128 rows take) for every batch
 *   "skip_finished"  (1) rows that emitted eos_id stop streaming their KV cache (their later tokens are not meaningful);
 *                        0 = every row keeps decoding until all rows have stopped, like the reference loop
 *   "share_keys"     (1) with skip_finished, batches of 128 rows: once two thirds of the rows have finished (seen by the
 *                        16-step stop poll), the decode-attention CTAs of finished rows take a share of the keys of the rows
 *                        that are still decoding (3 or 4 CTAs per row and kv head, work list kept on the device by the step
 *                        kernel, partial softmax states merged in a fixed order by the CTA that finishes last);
 *                        0 = finished rows only drop their own K/V stream
 *   "o_tail", "down_tail" (432, 848) shape of the decode tail GEMMs, cluster size * 100 + tile columns; other shapes
 *                        exist in lab builds only (measured: profiles/r2_decode_ab_tail_shapes.jsonl)
 *   "prefill_attn"   (1) causal prefill attention: 1 = tcgen05 kernel (TMA-fed bf16 operand planes, S / O in TMEM);
 *                        0 = the mma.sync kernel of round 1, lab builds only (MB_BUILD_LAB=1; an error otherwise)
 *   "attn_variant"   (2) decode attention kernel: 2 = warp-autonomous (each warp streams its own 16-key chunks with one bulk
 *                        copy per chunk and operand, no block barrier in the loop); lab builds only: 1 = the same with
 *                        16-byte cp.async pieces, 0 = the 64-key tile kernel of round 1 (an error otherwise)
 *   "kv_prefetch"    (0) lab tile kernel only: keys per (row, kv head) stream prefetched into L2 while the kernel waits
 *                        for its predecessor (-1 = the whole immutable history); measured slower, off
 *   "decode_tails"   (1) 1 = o_proj / down_proj of a decode layer as cluster split-K GEMMs that finish the residual add
 *                        and the (deferred) RMSNorm themselves: 5 kernels per layer instead of 7
 *   "decode_cluster" (0) lab builds only: gate/up (+SwiGLU) and QKV (+RoPE, KV write) of a decode layer as cluster split-K
 *                        GEMMs (3 K slices per cluster); measured slower (two waves of clusters), ignored otherwise
 *   "cta_pairs"      (1) persistent GEMM (prefill / encoder, M >= 1024): 192- / 256-column tiles as cta_group::2 pairs (two
 *                        CTAs of a cluster on one TPC share a 256-row tile, each loading half of the weight tile);
 *                        0 = single CTAs (bit-identical results, ~9 % slower LM prefill)
 *   "epilogue_rows"  (0) persistent GEMM: 1 = the row-per-thread global accesses of round 1 in the plain / QKV epilogue
 *                        instead of the staged 128-byte row segments (A/B switch)
 *   "wide_tiles"     (-1) decode split-K GEMM tiling: 1 = 32-column tiles x 3 / 8 K slices, 0 = 16-column tiles x 3 / 4,
 *                        -1 = by policy (wide except MB_POLICY_FAST)
 *   "gemm_engine"    (1) 0 = mma.sync cross-check engine (lab builds only, MB_BUILD_LAB=1) */
int mb_set_option(void* h, const char* name, int value);
/* profiling aid: dev_trace_buf = {u32 n; u32 cap; {u64 globaltimer_ns; u32 id*16+phase; u32 smid} ev[cap][16]} in device
 * memory, zero-initialised, n = records used, cap = records available (NULL = off).  The first and the last CTA of every
 * decode-step kernel claim one 16-slot record at entry and stamp phases into it with plain stores: 0 entry, 1 return of
 * the programmatic-dependency wait, 2 exit (first CTA), 3 exit (last CTA), 4.. kernel-specific (see gemm_skinny.cu,
 * lm.cu), 15 entry of the last CTA; id = kind*1000 + group*100 + layer (kind 1 QKV, 2 attention, 3 o_proj, 4 add+norm,
 * 5 gate/up, 6 down, 7 add+norm, 8 lm_head).  tools/decode_timeline.py prints the timeline. */
int mb_set_trace(void* h, void* dev_trace_buf);
long long mb_kernel_launches(void* h);       /* kernels launched by this handle so far (counting graph replays) */

/* ---- stages (each can be profiled / parity-checked in isolation) ---- */
/* wave [n_clips,320000] f32 -> logmel_out [n_clips,1001,64] (pre-BN, may be NULL), bn_out (post-bn0, may be NULL) */
int mb_frontend(void* h, const float* wave, int n_clips, float* logmel_out, float* bn_out, void* stream);
/* wave1, wave2 [B,320000] -> rows_out [2][B][33][576] f32 (may be NULL; the handle keeps its own copy for mb_prefix) */
int mb_encode(void* h, const float* wave1, const float* wave2, int B, float* rows_out, void* stream);
/* SURVEY 8 row f4 -- the encoder outputs the reference returns as od1/od2 (mellow/model/mellow.py:100-108,
 * htsat.py:782-796,950-955) for the n_clips = 2*B clips of the last mb_encode / mb_generate (audio1 rows first):
 * clipwise_out [n,527] = sigmoid(mean_t conv), framewise_rows_out [n,32,527] = sigmoid(conv) (the reference repeats
 * each row 32x to 1024 frames), latent_out [n,768], frame_embed_out [n,32,768] = c2l(frame rows) (rows 1.. of the
 * reference's `embedding`, before the 32x repeat).  Any pointer may be NULL. */
int mb_encode_heads(void* h, int n_clips, float* clipwise_out, float* framewise_rows_out, float* latent_out,
                    float* frame_embed_out, void* stream);
/* debug tap: run the encoder on `wave` [n_clips,320000] and copy the residual stream after stage `stage`:
 * 0 = patch embed [n,4096,96], 1..4 = after Swin stage (incl. merging) [n,1024,192] [n,256,384] [n,64,768] [n,64,768],
 * 5 = latent [n,768] followed by c2l frame rows [n,32,768] */
int mb_encode_tap(void* h, const float* wave, int n_clips, int stage, float* out, void* stream);
/* input_ids [B,129] i32 -> prefix_out [B,389,576] f32 (may be NULL); uses the rows of the last mb_encode */
int mb_prefix(void* h, const int* input_ids, int B, float* prefix_out, void* stream);
/* alternative to mb_encode+mb_prefix for tests: load an externally built prefix [B,389,576] */
int mb_set_prefix(void* h, const float* prefix, int B, void* stream);
/* LM prefill over the 389-token prefix, fills the KV cache; logits_out [B,49152] f32 of the last position (may be NULL) */
int mb_prefill(void* h, int B, float* logits_out, void* stream);
/* The reference's inner seams used by _generate_batch (wrapper.py:217-218, 237):
 * mb_lm_forward_last = model.caption_decoder.lm(inputs_embeds=embeds).logits[:, -1]: one cache-less causal forward over
 * embeds [B,S,576] f32 (1 <= S <= 389 + max_new_tokens, B*S <= max_batch*389) -> logits_out [B,49152] f32; it
 * overwrites the prefix / KV state of the handle.  mb_embed_tokens = lm.model.embed_tokens: ids [n] i32 -> out [n,576]. */
int mb_lm_forward_last(void* h, const float* embeds, int B, int S, float* logits_out, void* stream);
int mb_embed_tokens(void* h, const int* ids, int n, float* out, void* stream);
/* decode loop.  tokens_out [B,max_len] i32 (row stride max_len); *steps_out_host = number of valid columns (the
 * reference breaks once every row has emitted eos_id).  Each row is valid up to and including its first eos_id: the
 * reference discards what follows (wrapper.py:254), and finished rows stop streaming their KV cache here (option "skip_finished").  logits_dump [max_len][B][49152] f32 (NULL = off);
 * forced_tokens [B,max_len] i32 (NULL = off): teacher forcing, tokens_out still records the model's own argmax. */
int mb_decode(void* h, int B, int max_len, float temperature, float top_p, int eos_id, int* tokens_out,
              int* steps_out_host, float* logits_dump, const int* forced_tokens, void* stream);
/* whole path from device buffers */
int mb_generate(void* h, const float* wave1, const float* wave2, const int* input_ids, int B, int max_len,
                float temperature, float top_p, int eos_id, int* tokens_out, int* steps_out_host, void* stream);
/* whole path from HOST buffers (pinned recommended): H2D of the inputs and D2H of the tokens are inside the call */
int mb_generate_host(void* h, const float* wave1_host, const float* wave2_host, const int* input_ids_host, int B,
                     int max_len, float temperature, float top_p, int eos_id, int* tokens_out_host,
                     int* steps_out_host, void* stream);

/* ---- audio ingest (replaces torchaudio Resample + the tile/crop of wrapper.py:146-167) ----
 * pcm [n_in] f32 at orig*g Hz -> out [n_out] f32 at new*g Hz, n_out = ceil(new*n_in/orig); kernel [new][klen] f32 is the
 * windowed-sinc filter bank (klen = 2*width + orig) built by the host exactly like torchaudio's. */
int mb_audio_resample(void* h, const float* pcm, long long n_in, int orig, int new_, const float* kernel, int klen,
                      int width, float* out, long long n_out, void* stream);
/* out [320000] = samples tiled from the start (total < 320000, start = 0) or cropped at `start` (total >= 320000) */
int mb_audio_fit(void* h, const float* samples, long long total, long long start, float* out, void* stream);

/* ---- op-level test hooks (parity tests of single kernels) ---- */
/* C[M,N] = A[M,K] * W[N,K]^T (+bias) with the library's GEMM engine and operand policy; fp32 device in/out */
int mb_op_gemm(void* h, const float* A, const float* W, const float* bias, float* C, int M, int N, int K, int act,
               void* stream);

/* bench hook: launches the decode-attention kernel (+ its split combine) `iters` times over the handle's KV cache at
 * context length `ctx`, cycling through the 30 layer caches so no launch re-reads L2-resident data; the caller
 * brackets the call with CUDA events on `stream`. */
int mb_bench_decode_attention(void* h, int B, int ctx, int iters, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MELLOW_B200_H */
