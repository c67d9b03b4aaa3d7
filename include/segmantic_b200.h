/*
 * segmantic_b200 -- C ABI of the B200-native volumetric prediction path of dyollb/segmantic.
 *
 * The reference is pure Python: it has no FFI.  Its "operator interface" for this path is the set
 * of third-party calls it composes; each entry point below replaces one of them and cites the
 * reference call site (paths relative to /root/reference/src/segmantic):
 *
 *   sgm_unet_create / sgm_unet_forward   <- Net.__init__ + Net.forward        seg/monai_unet.py:99-124, 221-222
 *                                           (monai.networks.nets.UNet, eval mode, BN folded)
 *   sgm_sw_accumulate / sgm_sw_finalize  <- SlidingWindowInferer(...)(img,net) seg/monai_unet.py:637-639, 665
 *                                           + AsDiscreted(argmax=True)         seg/monai_unet.py:622
 *   sgm_resample_trilinear[_argmax]      <- Spacingd / Invertd(Spacingd)       seg/monai_unet.py:173-174, 615-621
 *   sgm_resample_itk                     <- sitk.ResampleImageFilter           image/processing.py:60-70, 87-97
 *   sgm_normalize_intensity / sgm_foreground_bbox
 *                                        <- NormalizeIntensityd/CropForegroundd seg/monai_unet.py:163-169
 *   sgm_ensemble_mean_argmax / _vote / _select_best
 *                                        <- MeanEnsembled / VoteEnsembled / SelectBestEnsembled + AsDiscreted
 *                                                                               seg/monai_unet.py:848-1004, seg/transforms.py:15-61
 *   sgm_confusion_matrix                 <- confusion_matrix + DiceMetric / ConfusionMatrixMetric inputs
 *                                                                               seg/evaluation.py:96-125, seg/monai_unet.py:640-725
 *
 * Conventions: every pointer named *_dev is a device pointer on the CURRENT CUDA device; the caller
 * (PyTorch in the Python host) owns all device buffers including the workspace; the library owns
 * only the packed weights inside an sgm_unet handle.  All calls are asynchronous on `stream`
 * (a cudaStream_t passed as void*).  Volumes are planar float32 [C][D0][D1][D2] with D2 fastest --
 * exactly a contiguous MONAI/torch tensor [C, X, Y, Z].  Return value 0 = OK, negative = error
 * (sgm_last_error() gives the thread-local message).  No exceptions cross the boundary.  There is no
 * CPU fallback: without a CUDA device every compute entry point returns SGM_ERR_CUDA.
 */
#ifndef SEGMANTIC_B200_H
#define SEGMANTIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SGM_API __attribute__((visibility("default")))
#else
#define SGM_API
#endif

#define SGM_OK 0
#define SGM_ERR_INVALID (-1)
#define SGM_ERR_CUDA (-2)
#define SGM_ERR_WORKSPACE (-3)
#define SGM_ERR_UNSUPPORTED (-4)

#define SGM_KIND_CONV 0
#define SGM_KIND_CONV_TRANSPOSE 1
#define SGM_KIND_IDENTITY 2

#define SGM_PRECISION_FP32 0 /* CUDA-core FFMA path, fp32 storage; verification / reference default */
#define SGM_PRECISION_BF16 1 /* tcgen05 implicit-GEMM path, bf16 storage, fp32 accumulation in TMEM */

#define SGM_MAX_LEVELS 8
#define SGM_MAX_STARTS 128

/* One convolution of the UNet, BatchNorm already folded (float64 fold on the host). */
typedef struct sgm_conv_desc {
  int32_t kind;        /* SGM_KIND_* */
  int32_t cin, cout;
  int32_t kernel;      /* 3 or 1 along every used spatial axis */
  int32_t stride;      /* 1 or 2 */
  int32_t has_act;     /* PReLU (single slope) follows */
  float alpha;         /* PReLU slope */
  const float* weight; /* HOST fp32, torch layout: conv [O][I][k..], transposed conv [I][O][k..] */
  const float* bias;   /* HOST fp32 [O] */
} sgm_conv_desc;

/* MONAI UNet(num_res_units=2) as built at seg/monai_unet.py:114-124.  `convs` lists the
 * convolutions in canonical order: for each down level (unit0, unit1, residual), bottom (unit0,
 * unit1, residual), then for each up level deepest first (transposed conv, residual-unit conv). */
typedef struct sgm_unet_desc {
  int32_t spatial_dims; /* 2 or 3 (config key `spatial_dims`) */
  int32_t in_channels;  /* config key `num_channels` */
  int32_t out_channels; /* num_classes = len(tissue_list) */
  int32_t n_levels;     /* len(channels) - 1 */
  int32_t channels[SGM_MAX_LEVELS];
  int32_t strides[SGM_MAX_LEVELS];
  int32_t precision; /* SGM_PRECISION_* */
  int32_t n_convs;
  const sgm_conv_desc* convs;
} sgm_unet_desc;

typedef struct sgm_unet sgm_unet;

/* Sliding-window schedule (MONAI sliding_window_inference).  Windows are the
 * cartesian product of the per-axis starts, axis 0 slowest.  The importance map of a window voxel
 * (i,j,k) is max((imap0[i]*imap1[j])*imap2[k], imap_floor); "constant" mode = tables of ones, floor 0.
 * For 2-D networks dims[0] = roi[0] = 1.
 *
 * Slab execution (multi-GPU): a rank runs the windows whose axis-0 start index lies in
 * [a0_begin, a0_end) and accumulates only planes [acc_x0, acc_x0+acc_nx) of axis 0; `vol_dev`
 * holds planes [vol_x0, vol_x0+vol_nx).  Single GPU: a0 range = all, x0 = 0, nx = dims[0]. */
typedef struct sgm_sw_cfg {
  int32_t dims[3]; /* full (padded) volume extent the schedule was computed for */
  int32_t roi[3];
  int32_t n_starts[3];
  int32_t starts[3][SGM_MAX_STARTS];
  int32_t sw_batch;         /* windows per network launch (results do not depend on it) */
  const float* imap[3];     /* HOST fp32 tables, length roi[d] */
  float imap_floor;
  int32_t a0_begin, a0_end; /* axis-0 start-index range executed by this call */
  int32_t vol_x0, vol_nx;   /* axis-0 planes present in vol_dev */
  int32_t acc_x0, acc_nx;   /* axis-0 planes present in acc_dev / outputs */
} sgm_sw_cfg;

SGM_API const char* sgm_last_error(void);
SGM_API int32_t sgm_version(void);

/* Net.__init__ (seg/monai_unet.py:99-124): packs + uploads the folded weights to the current device. */
SGM_API int32_t sgm_unet_create(const sgm_unet_desc* desc, sgm_unet** out);
SGM_API void sgm_unet_destroy(sgm_unet* net);
/* Bytes of caller-owned scratch needed to run `batch` ROI windows through the network. */
SGM_API int64_t sgm_unet_workspace_bytes(const sgm_unet* net, const int32_t roi[3], int32_t batch);
/* Net.forward (seg/monai_unet.py:221-222): x_dev [B][Cin][roi] fp32 -> logits_dev [B][C][roi] fp32. */
SGM_API int32_t sgm_unet_forward(sgm_unet* net, const float* x_dev, float* logits_dev, int32_t batch,
                         const int32_t roi[3], void* workspace_dev, int64_t workspace_bytes,
                         void* stream);
/* Synchronises `stream` and reports a tcgen05 pipeline timeout raised by any earlier launch on this
 * handle (the kernels bound every mbarrier wait instead of hanging the GPU). */
SGM_API int32_t sgm_unet_check(sgm_unet* net, void* stream);
/* Diagnostic: run convolution `conv_index` (canonical order) alone on caller-provided CG8 tensors
 * ([n][cg][d0][d1][d2][8], fp32 or bf16 per the handle's precision), on the CUDA-core family
 * (use_tc = 0) or the tcgen05 family (use_tc = 1; fused = 1 runs a strided down block's unit0 together
 * with its residual-branch conv into out / out2).  Writes the output extent to out_dims. */
SGM_API int32_t sgm_debug_conv(sgm_unet* net, int32_t conv_index, int32_t use_tc, int32_t fused,
                               const void* in0, int32_t cg0, const void* in1, int32_t cg1, const void* res,
                               void* out, void* out2, int32_t n, const int32_t in_dims[3],
                               int32_t out_dims[3], void* stream);
/* Per-convolution device timing: while on, every conv launch is bracketed by a CUDA-event pair on the
 * launching stream.  sgm_unet_get_profile synchronises, returns the accumulated milliseconds and launch
 * counts per convolution (canonical order) plus one last slot for the gather-blend kernel
 * (n = n_convs + 1) and resets the counters. */
SGM_API int32_t sgm_unet_set_profiling(sgm_unet* net, int32_t on);
SGM_API int32_t sgm_unet_get_profile(sgm_unet* net, double* ms, int64_t* launches, int32_t n, void* stream);
/* Number of kernel launches the last forward / sw_accumulate on this handle enqueued. */
SGM_API int64_t sgm_unet_last_launch_count(const sgm_unet* net);

/* inferer(inputs, network) (seg/monai_unet.py:665) without the final division: adds every window's
 * importance-weighted logits to acc_dev [C][acc_nx][dims1][dims2] (zero it first), in MONAI's
 * window order, reading windows straight from vol_dev [Cin][vol_nx][dims1][dims2]. */
SGM_API int32_t sgm_sw_accumulate(sgm_unet* net, const float* vol_dev, const sgm_sw_cfg* cfg, float* acc_dev,
                          void* workspace_dev, int64_t workspace_bytes, void* stream);
SGM_API int64_t sgm_sw_workspace_bytes(const sgm_unet* net, const sgm_sw_cfg* cfg);
/* `out / count` (count map recomputed analytically in MONAI's accumulation order) and, optionally,
 * AsDiscreted(argmax=True) (seg/monai_unet.py:622; ties -> lowest class) and channel softmax.
 * Any of logits_dev [C][..] / labels_dev (uint8 [..]) / probs_dev [C][..] may be NULL. */
SGM_API int32_t sgm_sw_finalize(const float* acc_dev, int32_t channels, const sgm_sw_cfg* cfg,
                        float* logits_dev, uint8_t* labels_dev, float* probs_dev, void* stream);

/* The whole inferer in one call with a DEFERRED blend: every window's importance-weighted logits are
 * stored once in the workspace ([window][C][roi] fp32) and a single streaming kernel then sums, for each
 * output voxel, its covering windows in MONAI's window order, divides by the count and takes
 * argmax / softmax.  Bit-identical to sgm_sw_accumulate + sgm_sw_finalize (same fp32 operations in
 * the same order) without any read-modify-write; needs windows * C * roi voxels * 4 extra workspace
 * bytes.  Any of logits_dev / labels_dev / probs_dev may be NULL. */
SGM_API int64_t sgm_sw_predict_workspace_bytes(const sgm_unet* net, const sgm_sw_cfg* cfg);
SGM_API int32_t sgm_sw_predict(sgm_unet* net, const float* vol_dev, const sgm_sw_cfg* cfg, float* logits_dev,
                               uint8_t* labels_dev, float* probs_dev, void* workspace_dev,
                               int64_t workspace_bytes, void* stream);

/* The two halves of sgm_sw_predict as separate calls, for the multi-GPU driver with window OWNERSHIP: a rank
 * computes the importance-weighted logits of a contiguous range of the schedule's windows (MONAI order, axis 0
 * slowest) into wl_dev ([w_count][C][roi] fp32), receives the windows of the previous rank that cover its output
 * planes over NVLink, and blends.  sgm_sw_windows reads cfg->vol_x0 / vol_nx (the input planes present) and
 * cfg->sw_batch; sgm_sw_blend reads cfg->a0_begin / a0_end (the window rows held in wl_dev, which starts at the
 * first window of row a0_begin) and cfg->acc_x0 / acc_nx (the output planes).  Same arithmetic as sgm_sw_predict:
 * bit-identical to the single-device result.  sgm_sw_blend needs 16 KiB of device scratch. */
SGM_API int64_t sgm_sw_windows_workspace_bytes(const sgm_unet* net, const sgm_sw_cfg* cfg);
SGM_API int32_t sgm_sw_windows(sgm_unet* net, const float* vol_dev, const sgm_sw_cfg* cfg, int64_t w_first,
                               int64_t w_count, float* wl_dev, void* workspace_dev, int64_t workspace_bytes,
                               void* stream);
SGM_API int32_t sgm_sw_blend(const sgm_sw_cfg* cfg, int32_t channels, const float* wl_dev, float* logits_dev,
                             uint8_t* labels_dev, float* probs_dev, void* scratch_dev, void* stream);

/* Spacingd / its inverse (seg/monai_unet.py:173-174,615-621): out[c][o] = trilinear(in[c], A*o + t),
 * grid_sample(bilinear, padding border, align_corners False) semantics, float64 arithmetic.
 * xform = row-major 3x4 [A|t] mapping OUTPUT voxel indices to INPUT voxel indices. */
SGM_API int32_t sgm_resample_trilinear(const float* in_dev, const int32_t in_dims[3], int32_t channels,
                               float* out_dev, const int32_t out_dims[3], const double xform[12],
                               void* stream);
/* Invertd(Spacingd) on C-channel logits fused with AsDiscreted(argmax): never materialises the
 * resampled logits.  Output voxels are written at out_dev[(o0+pad_lo[0])..] of a [full_dims] uint8
 * volume that the caller pre-fills with 0 (inverse CropForeground = zero padding -> label 0). */
SGM_API int32_t sgm_resample_trilinear_argmax(const float* in_dev, const int32_t in_dims[3], int32_t channels,
                                      uint8_t* out_dev, const int32_t out_dims[3],
                                      const double xform[12], void* stream);

/* sitk.ResampleImageFilter with identity transform (image/processing.py:60-70,87-97).  Arrays are in
 * ITK index order, x fastest: element (x,y,z) at x + nx*(y + ny*z); dims = {nx,ny,nz} (nz = 1 for 2-D).
 * dtype: 0 uint8, 1 int16, 2 uint16, 3 float32, 4 int32.  Geometry in float64: index->physical of the
 * output grid (row-major 3x3 + origin) and physical->index of the input (3x3 + origin).  Outside the
 * input buffer ([-0.5, n-0.5) per axis) -> default_value.  nearest: floor(c+0.5); linear: ITK's
 * clamped-neighbour lerp in double, cast with clamping + truncation to the pixel type. */
SGM_API int32_t sgm_resample_itk(const void* in_dev, int32_t dtype, const int32_t in_dims[3], void* out_dev,
                         const int32_t out_dims[3], const double out_index_to_phys[9],
                         const double out_origin[3], const double in_phys_to_index[9],
                         const double in_origin[3], int32_t nearest, double default_value,
                         void* stream);

/* NormalizeIntensityd(nonzero=False, channel_wise=True) (seg/monai_unet.py:163): per channel
 * (x-mean)/std, biased std, std==0 -> 1.  In place allowed.  scratch_dev: >= 4096 doubles. */
SGM_API int32_t sgm_normalize_intensity(const float* in_dev, float* out_dev, int32_t channels, int64_t voxels,
                                double* scratch_dev, void* stream);
/* CropForegroundd(select_fn = x > 0) (seg/monai_unet.py:164-169): bbox over any channel.
 * bbox_dev: int32[6] = {lo0,lo1,lo2,hi0,hi1,hi2} (hi exclusive; empty -> all zeros). */
SGM_API int32_t sgm_foreground_bbox(const float* in_dev, int32_t channels, const int32_t dims[3],
                            int32_t* bbox_dev, void* stream);

/* Evaluation branch of predict() (seg/monai_unet.py:672-704; seg/evaluation.py:96-125): confusion matrix of two uint8
 * label maps, cm[t * num_classes + p] = #{voxels with true label t and predicted label p} (rows = y, columns =
 * y_pred: the convention of sklearn.metrics.confusion_matrix the reference's docstring names).  Pairs with a label
 * >= num_classes are not counted; their number is written to ignored_dev (may be NULL).  num_classes <= 64. */
SGM_API int32_t sgm_confusion_matrix(const uint8_t* y_pred_dev, const uint8_t* y_dev, int64_t voxels,
                                     int32_t num_classes, int64_t* cm_dev, int64_t* ignored_dev, void* stream);

/* Ensemble combination (seg/monai_unet.py:848-1004, seg/transforms.py:15-61), voxel-wise over stacked model outputs:
 *   mean:        logits [E][C][V] float32 -> mean_m(x_m * w_m / mean(w)) (MONAI MeanEnsemble; weights NULL = equal; host
 *                array) -> argmax labels (ties -> lowest class); mean_dev (may be NULL) receives the [C][V] mean.
 *   vote:        labels [E][V] uint8 -> majority class (MONAI VoteEnsemble(num_classes): ties -> lowest class).
 *   select_best: labels [E][V] uint8 + (tissue id, model index) pairs in dictionary order -> voxels the chosen model
 *                labels `tissue` get `tissue`, later pairs overwrite earlier ones, unclaimed voxels are 0. */
SGM_API int32_t sgm_ensemble_mean_argmax(const float* logits_dev, int32_t n_models, int32_t num_classes, int64_t voxels,
                                         const float* weights, uint8_t* labels_dev, float* mean_dev, void* stream);
SGM_API int32_t sgm_ensemble_vote(const uint8_t* labels_in_dev, int32_t n_models, int32_t num_classes, int64_t voxels,
                                  uint8_t* labels_out_dev, void* stream);
SGM_API int32_t sgm_ensemble_select_best(const uint8_t* labels_in_dev, int32_t n_models, int64_t voxels,
                                         const int32_t* tissue_ids, const int32_t* model_ids, int32_t n_pairs,
                                         uint8_t* labels_out_dev, void* stream);

/* NVLink peer-memory exchange of the multi-GPU driver (one process per GPU; no counterpart in the reference, which
 * predicts on gpu_ids[0] only: seg/utils.py:4-12 -- this is the data-parallel driver BASELINE.json's north_star adds).
 * A rank exports a device buffer (any pointer inside a cudaMalloc allocation, e.g. a torch tensor of the default
 * caching allocator) as a CUDA IPC handle; the neighbouring process maps it (sgm_p2p_open returns the BASE of the
 * allocation, add `offset`), pushes data with the copy engines (sgm_p2p_put: cudaMemcpyAsync device-to-device over
 * NVLink, no SM involved) and raises a 32-bit counter in the peer's memory (sgm_p2p_signal, ordered after everything
 * queued before it on `stream`); the owner waits on its own stream until the counter reaches `value` (sgm_p2p_wait;
 * after timeout_s seconds it gives up and sets *timeout_flag_dev = 1 instead of hanging the device). */
SGM_API int32_t sgm_p2p_export(const void* ptr_dev, uint8_t handle[64], int64_t* offset);
SGM_API int32_t sgm_p2p_open(const uint8_t handle[64], void** base_out);
SGM_API int32_t sgm_p2p_close(void* base);
SGM_API int32_t sgm_p2p_put(void* dst_peer_dev, const void* src_dev, int64_t bytes, void* stream);
SGM_API int32_t sgm_p2p_signal(uint32_t* flag_peer_dev, uint32_t value, void* stream);
SGM_API int32_t sgm_p2p_wait(const uint32_t* flag_dev, uint32_t value, int32_t* timeout_flag_dev, double timeout_s,
                             void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SEGMANTIC_B200_H */
