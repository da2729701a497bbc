/*
 * dcb200 -- C ABI of the B200-native DynamiCrafter denoising hot path.
 *
 * Nothing like this exists in the reference (pure PyTorch): each entry point
 * replaces a class of ATen / cuDNN / cuBLAS call sites of the reference's hot path, cited
 * below as file:line under /root/reference.  The reference-side binding is a ctypes stub
 * (INTEGRATION.md); the Python mirror classes in dynamicrafter_b200/ are that binding.
 *
 * Conventions: plain pointers and sizes only; every pointer is DEVICE memory unless named
 * h_*; the library never allocates, never synchronises and never throws.  All functions
 * return 0 on success or a negative dcb_status; dcb_last_error() gives a thread-local
 * message.  `stream` is a cudaStream_t passed as void*.  Kernels are stateless, so calls are
 * CUDA-graph capturable and thread-safe when callers use distinct buffers.
 *
 * Activation layout everywhere: channels-last token matrix [rows, C] with rows ordered
 * (batch, frame, y, x); bf16 operands feed the tensor cores, the residual stream is fp32.
 */
#ifndef DCB200_H
#define DCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    DCB_OK = 0,
    DCB_ERR_INVALID = -1,   /* bad argument (shape / alignment / null pointer) */
    DCB_ERR_CUDA = -2,      /* a CUDA runtime / driver call failed */
    DCB_ERR_UNSUPPORTED = -3,
    DCB_ERR_DEVICE = -4     /* not an sm_100 device */
} dcb_status;

#define DCB_MAX_TAPS 9

const char* dcb_last_error(void);
int dcb_abi_version(void);
/* 0 if the current device is compute capability 10.x, DCB_ERR_DEVICE otherwise. */
int dcb_check_device(void);

/* ------------------------------------------------------------------------------------------
 * dcb_tap_gemm: tcgen05/TMEM/TMA implicit GEMM,  out[r, n] = epi( sum_{tap,c} A[r + off(tap), c] * W[n, tap*C + c] )
 *
 * Replaces: nn.Linear / 1x1 conv  (attention.py:53-57,75-76,269,290,336,418,438; ae_modules.py:31-50)
 *           nn.Conv2d 3x3 s1 p1   (openaimodel3d.py:154,179,185,96,386,545; ae_modules.py:117,161,173,491,533)
 *           nn.Conv3d (3,1,1)     (openaimodel3d.py:255-266)
 *
 * A is a bf16 tensor viewed as [D3a, D2, D1, C] (C contiguous; strides s1,s2,s3 in elements,
 * multiples of 8).  The output domain is [D3, D2, D1]; an M-tile is a box (b1,b2,b3) of it with
 * b1*b2*b3 <= 128.  For each tap the box is shifted by (o1,o2,o3); reads outside the tensor are
 * zero (TMA out-of-bounds fill) which implements the convolution padding.  W is bf16 [N, ntaps*C]
 * row-major ("K-major").  C % 64 == 0.
 *
 * Epilogue (fp32): acc + bias[n] + group_bias[r / rows_per_group, n]; if geglu, tile columns
 * [0,BN/2) are values and [BN/2,BN) gates and out = v * gelu_erf(g) (N_out = N/2, attention.py:421-422);
 * then + residual[r, n]; written to out_f32 and/or out_bf16 (row stride ld_out elements).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    const void* a;            /* bf16 */
    int64_t a_s1, a_s2, a_s3; /* element strides of dims 1..3 of A (dim 0 = channel, stride 1) */
    int32_t C;                /* channels per tap */
    int32_t D1, D2, D3;       /* output domain */
    int32_t A1, A2, A3;       /* extents of A's dims 1..3 (>= domain; larger for phase-plane tricks) */
    int32_t b1, b2, b3;       /* M-tile box */
    int32_t ntaps;
    int32_t taps[DCB_MAX_TAPS][3];
    const void* w;            /* bf16 [N, ntaps*C] */
    int32_t N;                /* GEMM N (rows of w) */
    int32_t block_n;          /* 128 or 160 (0 = choose) */
    const float* bias;        /* [N_out] or NULL (for geglu: [N], value half then gate half per tile) */
    const float* group_bias;  /* [groups, ld_group_bias] or NULL */
    int32_t rows_per_group;
    int32_t ld_group_bias;
    const float* residual;    /* [rows, ld_residual] fp32 or NULL */
    int32_t ld_residual;
    int32_t geglu;
    float* out_f32;           /* either may be NULL, not both */
    void* out_bf16;
    int32_t ld_out;
    float out_scale;          /* applied to acc before bias (1.0 = none) */
} dcb_tap_gemm_args;

int dcb_tap_gemm(const dcb_tap_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DCB200_H */
