/*
 * Constants of the canonical arithmetic shared by the CUDA kernels
 * (meshanything_b200/csrc) and the CPU oracle (oracle/decoder_oracle.c).
 *
 * Only numbers live here.  The two sides implement the arithmetic
 * independently (SIMT lanes on the GPU, scalar loops emulating the lanes on
 * the CPU) and must agree bit for bit; DESIGN.md section 3 states the order.
 */
#ifndef MA_CANON_CONSTANTS_H
#define MA_CANON_CONSTANTS_H

/* ma_exp(x) = 2^(x*log2e): n = rint(y), f = y-n in [-0.5,0.5],
 * 2^f ~= Horner(MA_EXP2_C6..C0) with fmaf, then exponent add of n.
 * Relative error 7.2e-8 on [-0.5,0.5].  Arguments below MA_EXP_FLUSH give 0. */
#define MA_LOG2E      1.44269504088896340736f
#define MA_EXP_FLUSH  (-80.0f)
#define MA_EXP2_C0    1.0f
#define MA_EXP2_C1    0x1.62e430p-1f
#define MA_EXP2_C2    0x1.ebfbe0p-3f
#define MA_EXP2_C3    0x1.c6af6cp-5f
#define MA_EXP2_C4    0x1.3b2a1cp-7f
#define MA_EXP2_C5    0x1.5f0896p-10f
#define MA_EXP2_C6    0x1.444004p-13f

/* LayerNorm epsilon (torch.nn.LayerNorm default, used by HF OPT/BERT and by
 * michelangelo's transformer_blocks.py:104,106). */
#define MA_LN_EPS     1e-5f

/* Attention: keys are processed in chunks of MA_ATTN_CHUNK positions; inside a
 * chunk position r goes to group-lane r % 32 (8 warps x 4 groups of 8 lanes). */
#define MA_ATTN_CHUNK 256
#define MA_HEAD_DIM   64

#endif
