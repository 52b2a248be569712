/*
 * lbad_cuda.h — internal C interface between the host-side API layer (plain C: LBAudioDetective.c,
 * LBAudioDetectiveFingerprint.c, LBAudioDetectiveDatabase.c) and the CUDA layer (lbad_extract.cu,
 * lbad_search.cu, lbad_synth.cu).  Plain pointers and sizes only.  Not installed; not part of the public ABI.
 */
#ifndef LBAD_CUDA_H
#define LBAD_CUDA_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define LBAD_ROWS_PER_FRAME 128u     /* LBAudioDetective.m:25 */
#define LBAD_MAX_BANDS      64u
#define LBAD_MAX_WINDOW     2048u
#define LBAD_MIN_WINDOW     256u
#define LBAD_MAX_SUBLEN     512u

#define LBAD_OK              0
#define LBAD_ERR_ARG         1       /* kLBAudioDetectiveArgumentInvalid */
#define LBAD_ERR_NODEVICE   (-7001)
#define LBAD_ERR_CUDA       (-7002)

/* Geometry + host-computed tables (the band table follows LBAudioDetective.m:361-383 in double precision). */
typedef struct {
    uint32_t window;                 /* N, power of two, 256..2048 */
    uint32_t stride;                 /* hop */
    uint32_t bands;                  /* B, power of two, 4..64 */
    uint32_t sublen;                 /* L Booleans kept per subfingerprint */
    uint32_t klow[LBAD_MAX_BANDS];   /* first FFT bin of each band (m:382) */
    uint32_t khigh[LBAD_MAX_BANDS];  /* one past the last bin (m:383) */
    float    divisor[LBAD_MAX_BANDS];/* (Float32)(indices[i+1]-indices[i]) (m:404) */
    float    pos_scale;              /* (Float32)(width/2) with width = N/2 (m:373, m:391) */
} lbadcu_geometry;

typedef struct lbadcu_plan lbadcu_plan;

/* 0 if a CUDA device is usable */
int  lbadcu_device_available(void);
int  lbadcu_device_count(void);
const char* lbadcu_last_error(void);          /* per thread */
void lbadcu_set_last_error(const char* msg);

/* make `device` current for the calling thread (device < 0: leave it) / put the previous one back */
int  lbadcu_push_device(int device, int* prev);
void lbadcu_pop_device(int prev);

/* the plan is bound to the device that is current when it is created */
int  lbadcu_plan_create(const lbadcu_geometry* g, lbadcu_plan** out);
int  lbadcu_plan_device(const lbadcu_plan* p);
void lbadcu_plan_destroy(lbadcu_plan* p);
int  lbadcu_plan_fused_supported(const lbadcu_plan* p);
void* lbadcu_plan_stream(lbadcu_plan* p);
uint64_t lbadcu_plan_launches(const lbadcu_plan* p);
/* which = 0: FFT + band-energy kernel (the dominant one), 1: Haar / select / pack kernel */
uint32_t lbadcu_plan_timing(lbadcu_plan* p, int which, int enable, int reset, double* total_ms);

/* Extraction over DEVICE-resident PCM.  Clip c starts at d_pcm + c*clip_stride and has clip_len samples; every
 * clip yields frames = ((clip_len - N)/hop)/128 subfingerprints.  d_words: [clip][frame][2*W].
 * mode: 0 = auto (fused when supported), 1 = force fused, 2 = force generic.  d_images / d_haar (optional stage
 * dumps): [clip][frame][128][B].  Enqueues on `stream` (NULL = the plan's stream); does not synchronise. */
int  lbadcu_extract_device(lbadcu_plan* p, const float* d_pcm, uint32_t n_clips, uint64_t clip_len, uint64_t clip_stride,
                           uint32_t* d_words, float* d_images, float* d_haar, int mode, void* stream);
/* Host-memory front end: uploads in chunks on two streams, runs the kernels, downloads the packed words. */
int  lbadcu_extract_host(lbadcu_plan* p, const float* h_pcm, uint32_t n_clips, uint64_t clip_len, uint64_t clip_stride,
                         uint32_t* h_words, float* h_images, float* h_haar, int mode);
/* One batch shared by several plans (one per GPU), each called from its own host thread with the WHOLE batch and the same cursor
 * (TWO 64-bit words the caller sets to 0: the clips handed out so far, and the chunk size the first plan to arrive fixes for all): a plan takes its next chunk from the cursor when one of its three buffers has drained, so the
 * batch divides itself by how fast each GPU is fed.  Every clip is processed by exactly one plan; h_words is filled in clip order. */
int  lbadcu_extract_host_shared(lbadcu_plan* p, const float* h_pcm, uint32_t n_clips, uint64_t clip_len, uint64_t clip_stride, uint32_t* h_words, uint64_t* cursor);
/* same for signed 16-bit PCM (uploaded as 2 bytes per sample, converted on the device as x / 32768) */
int  lbadcu_extract_host_i16(lbadcu_plan* p, const int16_t* h_pcm, uint32_t n_clips, uint64_t clip_len, uint64_t clip_stride, uint32_t* h_words);
/* Haar + top-t + pack on device images [count][128][B] (generic kernel). */
int  lbadcu_transform_images_host(lbadcu_plan* p, const float* h_images, uint32_t count, float* h_haar, uint32_t* h_words);

/* ---- the Frame API's two computing functions on frames of any shape (lbad_frame.cu; current device) ---- */
/* h_a: [rows][cols] in host memory, transformed in place (LBAudioDetectiveFrame.m:113-153) */
int  lbadcu_frame_decompose_host(float* h_a, uint32_t rows, uint32_t cols);
/* h_a: n coefficients in flat order; h_out: 2t bytes, 1 where LBAudioDetectiveFrame.m:182-190 sets TRUE, 0 elsewhere */
int  lbadcu_frame_extract_host(const float* h_a, uint32_t n, uint32_t t, unsigned char* h_out);

/* ---- recording-rate -> processing-rate conversion (include/LBAudioDetectiveResample.h) ---- */
#define LBAD_RS_PHASES 64u
typedef struct {
    double in_rate, out_rate, rho2;  /* rho2 = in_rate / (out_rate * D) */
    uint32_t D, H1, T1;              /* integer pre-decimation; stage-1 half length and taps (T1 = 2 H1 + 1; unused when D = 1) */
    uint32_t H2, T2;                 /* stage-2 half length and taps (T2 = 2 H2) */
    float* g;                        /* [T1] */
    float* hc;                       /* [LBAD_RS_PHASES + 1][T2] */
} lbadcu_resample_design;
/* host-side filter design (lbad_resample_design.c) */
int  lbad_resample_design_create(double in_rate, double out_rate, lbadcu_resample_design* d);
void lbad_resample_design_free(lbadcu_resample_design* d);
uint64_t lbad_resample_out_len(double in_rate, double out_rate, uint64_t n_in);
typedef struct lbadcu_resampler lbadcu_resampler;
int  lbadcu_resampler_create(const lbadcu_resample_design* d, lbadcu_resampler** out);
void lbadcu_resampler_destroy(lbadcu_resampler* r);
uint64_t lbadcu_resampler_launches(const lbadcu_resampler* r);
/* clips of in_len recorded samples, in_stride apart -> out_len samples each, out_stride apart (device pointers) */
int  lbadcu_resample_device(lbadcu_resampler* r, const float* d_in, uint32_t n_clips, uint64_t in_len, uint64_t in_stride,
                            float* d_out, uint64_t out_len, uint64_t out_stride, void* stream);
int  lbadcu_resample_host(lbadcu_resampler* r, const float* h_in, uint64_t n_in, float* h_out, uint64_t n_out);
/* one recorded clip from host memory: upload, convert, extract, download n_words packed words — nothing but the words returns to the host */
int  lbadcu_process_recorded_host(lbadcu_plan* p, lbadcu_resampler* r, const float* h_in, uint64_t n_in, uint64_t out_len, uint32_t* h_words, size_t n_words);
int  lbadcu_device_alloc_floats(uint64_t n, float** out);
void lbadcu_device_free(void* p);

/* ---- search ---- */
typedef struct lbadcu_db lbadcu_db;
/* pairs_full = ceil(L/2): pairs every stored subfingerprint carries */
int  lbadcu_db_create(uint32_t words_per_plane, uint32_t pairs_full, lbadcu_db** out);
void lbadcu_db_destroy(lbadcu_db* db);
uint32_t lbadcu_db_clips(const lbadcu_db* db);
uint64_t lbadcu_db_subfps(const lbadcu_db* db);
uint32_t lbadcu_db_min_count(const lbadcu_db* db);
uint32_t lbadcu_db_max_count(const lbadcu_db* db);
void lbadcu_db_set_base(lbadcu_db* db, uint32_t base);
void* lbadcu_db_stream(lbadcu_db* db);
uint64_t lbadcu_db_launches(const lbadcu_db* db);
uint32_t lbadcu_db_timing(lbadcu_db* db, int enable, int reset, double* total_ms);
/* counts == NULL -> uniform_count for every clip; words on host or device.  producer_stream (device words only, may be NULL): the stream
 * the words are being produced on — the copy is ordered after the work enqueued there so far.  first_global_id >= 0: the clips carry
 * the global ids first_global_id, first_global_id + 1, ... (a shard whose clips are not one contiguous range); -1: id = base + index. */
int  lbadcu_db_append(lbadcu_db* db, const uint32_t* words, int words_on_device, uint32_t n_clips, const uint32_t* counts, uint32_t uniform_count,
                      void* producer_stream, int64_t first_global_id);
int  lbadcu_db_download(lbadcu_db* db, uint32_t* h_words, uint32_t* h_counts);
/* pairs = number of (P,M) bit pairs compared = ceil(min(range, L)/2).  Outputs [q][k]; d_all optional [q][clips]. */
int  lbadcu_db_search_device(lbadcu_db* db, const uint32_t* d_qwords, uint32_t n_q, uint32_t q_count, uint32_t pairs, uint32_t k,
                             float* d_scores, uint32_t* d_idx, float* d_all, void* stream);
int  lbadcu_db_search_host(lbadcu_db* db, const uint32_t* h_qwords, uint32_t n_q, uint32_t q_count, uint32_t pairs, uint32_t k,
                           float* h_scores, uint32_t* h_idx, float* h_all);
uint64_t lbadcu_db_compares_per_query(const lbadcu_db* db, uint32_t q_count);
/* merges [list][q][k] top-k lists (host memory) with the device merge kernel; order (score desc, index asc) */
int  lbadcu_merge_topk_host(const float* h_sc, const uint32_t* h_id, uint32_t n_lists, uint32_t n_q, uint32_t k, float* o_sc, uint32_t* o_id);

/* ---- a database sharded over several devices of this process (lbad_search.cu, "group") ---- */
typedef struct lbadcu_group lbadcu_group;
int  lbadcu_group_create(uint32_t words_per_plane, uint32_t pairs_full, const int* devices, uint32_t n_shards, lbadcu_group** out);
void lbadcu_group_destroy(lbadcu_group* g);
uint32_t lbadcu_group_shards(const lbadcu_group* g);
lbadcu_db* lbadcu_group_shard(lbadcu_group* g, uint32_t i);
int  lbadcu_group_shard_device(const lbadcu_group* g, uint32_t i);
uint64_t lbadcu_group_clips(const lbadcu_group* g);
uint64_t lbadcu_group_next_id(const lbadcu_group* g);
uint64_t lbadcu_group_launches(const lbadcu_group* g);
double lbadcu_group_last_search_ms(const lbadcu_group* g);
int  lbadcu_group_append(lbadcu_group* g, const uint32_t* h_words, uint32_t n_clips, const uint32_t* counts, uint32_t uniform_count);
int  lbadcu_group_append_one(lbadcu_group* g, const uint32_t* h_words, uint32_t count, uint64_t* out_id);
int  lbadcu_group_append_device(lbadcu_group* g, uint32_t shard, const uint32_t* d_words, uint32_t n_clips, uint32_t uniform_count, uint64_t first_global_id, void* producer_stream);
int  lbadcu_group_search_host(lbadcu_group* g, const uint32_t* h_qwords, uint32_t n_q, uint32_t q_count, uint32_t pairs, uint32_t k, float* h_scores, uint32_t* h_idx);

/* list l starts at d_sc + l * list_stride / d_id + l * list_stride (elements; 0 = n_q * k, lists back to back) */
int  lbadcu_merge_topk_device(const float* d_sc, const uint32_t* d_id, uint32_t n_lists, uint64_t list_stride, uint32_t n_q, uint32_t k, float* d_o_sc, uint32_t* d_o_id, void* stream);
/* one pair, LBAudioDetectiveFingerprintCompareToFingerprint(fp1, fp2) with pairs = ceil(min(range, L)/2); cached per-thread context */
int  lbadcu_compare_pair(uint32_t words_per_plane, uint32_t pairs, const uint32_t* w1, uint32_t c1, const uint32_t* w2, uint32_t c2, float* out);
/* the same on packed words that are already on the device; the score stays on the device (d_out) — no synchronisation */
int  lbadcu_compare_pair_device(uint32_t words_per_plane, uint32_t pairs, const uint32_t* d_w1, uint32_t c1, const uint32_t* d_w2, uint32_t c2, float* d_out, void* stream);
/* LBAudioDetectiveCompareAudioURLs as ONE device pipeline (LBAudioDetective.m:442-464): both clips uploaded, fingerprinted (as one
 * two-clip batch when their lengths agree) and compared without the words ever visiting the host; one synchronisation.  Both clips
 * must yield at least one subfingerprint. */
int  lbadcu_compare_pcm_host(lbadcu_plan* p, const float* h1, uint64_t n1, const float* h2, uint64_t n2, uint32_t pairs, float* out);

/* ---- synthetic PCM on the device (bench support; same formula as the host generator, device libm) ---- */
int  lbadcu_synth_device(float* d_out, uint32_t n_clips, uint64_t clip_len, uint64_t clip_stride, uint64_t first_clip_id,
                         uint64_t base_seed, double sample_rate, void* stream);
/* random rank-sign codes straight into packed words, for search timing at sizes extraction would take long to fill */
int  lbadcu_random_codes_device(uint32_t* d_words, uint64_t n_subfps, uint32_t words_per_plane, uint32_t pairs, uint64_t seed, uint64_t first_subfp, void* stream);

/* ---- microbenchmarks (roofline denominators measured on the box: FP32 FMA rate, POPC rate) ---- */
int  lbadcu_microbench(double* fp32_tflops, double* popc_gops, double* lop3_gops);

#ifdef __cplusplus
}
#endif
#endif
