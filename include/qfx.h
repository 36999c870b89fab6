/* qfx.h — C ABI of libqfx_b200.so: the B200 (sm_100a) replacement for the reference's per-step hot path
 *
 *     noise-predict forward (MMDiT) -> flow-matching MSE -> backward -> LoRA gradients
 *
 * The reference (tsiendragon/qwen-image-finetune, package `qflux`) is pure Python and has NO FFI of its own
 * (SURVEY.md §8b): its plug point is the `trainer.dit` nn.Module called from `BaseTrainer._compute_loss`
 * (/root/reference/src/qflux/trainer/base_trainer.py:473-476, qwen_image_edit_trainer.py:827-836).  Each entry
 * point below therefore cites the reference *Python* call it replaces.  All pointers are raw device pointers
 * (bf16 unless stated), all calls are asynchronous on `stream` (a cudaStream_t passed as void*), nothing here
 * synchronises with the host, nothing falls back to the CPU.  Return value: 0 on success, otherwise a
 * cudaError_t / negative argument-error code; `qfx_last_error()` returns a static message.
 *
 * INTEGRATION.md shows the ctypes binding the reference side would add.
 */
#ifndef QFX_H_
#define QFX_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* qfx_last_error(void);
int qfx_version(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Fused LoRA projection GEMM (tcgen05 + TMA).  Replaces one `peft.tuners.lora.Linear.forward` call
 * (injected at base_trainer.py:929-941; y = base(x) + lora_B(lora_A(x)) * alpha/r) or its autograd backward, for
 * up to QFX_MAX_PROBLEMS row-groups that share N, K (e.g. the image and text streams of one MMDiT block:
 * transformer_qwenimage.py:286-293).
 *
 *   trans_b = 0 :  out[M,N] = epi( alpha * ( A[M,K] . B[N,K]^T  +  A2[M,64*kb2] . B2[N,64*kb2]^T ) + bias )
 *   trans_b = 1 :  out[M,N] = epi( alpha * ( A[M,K] . B[K,N]    +  A2[M,64*kb2] . B2[64*kb2,N]   ) )      (dgrad)
 *
 * The LoRA factor pair (A2,B2) is appended to the K loop of the SAME tensor-core contraction (rank padded to a
 * multiple of 64 with zeros by the caller), so one accumulator in TMEM receives base + low-rank update.
 */
#define QFX_MAX_PROBLEMS 2

enum qfx_epilogue {
  QFX_EPI_BIAS = 0,       /* out = alpha*acc + bias                                                  (nn.Linear)            */
  QFX_EPI_GELU = 1,       /* out2 = u = acc + bias ; out = gelu_tanh(u)        (FeedForward "gelu-approximate" net.0)       */
  QFX_EPI_RESID_GATE = 2, /* out = resid + gate[row / rows_per_batch, :] * (acc + bias)   (transformer_qwenimage.py:473,480) */
  QFX_EPI_DGELU = 3       /* out = acc * gelu_tanh'(aux)                                   (autograd of net.0's GELU)        */
};

typedef struct {
  const void* A;   int64_t lda;  /* [M, K] row-major, lda in elements */
  const void* B;   int64_t ldb;  /* trans_b=0: [N, K];  trans_b=1: [K, N] */
  int M;
  const void* A2;  int64_t lda2; /* LoRA left factor  [M, >= a2_col0 + 64*kb2]  (NULL/kb2=0: none) */
  const void* B2;  int64_t ldb2; /* LoRA right factor trans_b=0: [N, 64*kb2];  trans_b=1: [64*kb2, N] */
  int kb2;                       /* number of appended 64-wide k-blocks */
  int a2_col0;                   /* first column of A2 to use */
  const void* bias;              /* [N] or NULL */
  void* out;       int64_t ldo;  /* [M, N] */
  void* out2;      int64_t ldo2; /* QFX_EPI_GELU: pre-activation u */
  const void* resid; int64_t ldr;/* QFX_EPI_RESID_GATE */
  const void* gate;  int64_t ldg; int rows_per_batch;
  const void* aux;   int64_t ldaux; /* QFX_EPI_DGELU: u */
} qfx_gemm_problem;

/* lora_group_n > 0 (trans_b = 0 only): output columns [g*lora_group_n, (g+1)*lora_group_n) use A2 columns
 * a2_col0 + g*64*kb2 .. (fused q|k|v projection with one LoRA pair per third).  block_n: 0 = auto, else 64/128/192/256. */
int qfx_gemm_bf16(const qfx_gemm_problem* probs, int nprob, int N, int K, int trans_b, int epilogue, float alpha,
                  int lora_group_n, int block_n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QFX_H_ */
