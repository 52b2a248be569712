/* TEST INFRASTRUCTURE — not part of the product (see oracle/README.md).
 * Plain-C restatement of the reference's fingerprint path; every function cites the reference lines it follows.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it. */
#ifndef LBAD_ORACLE_H
#define LBAD_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint32_t window;       /* N, LBAudioDetective.m:22 (2048)  */
    uint32_t stride;       /* hop, m:23 (64)                   */
    uint32_t bands;        /* pitch steps B, m:24 (32)         */
    uint32_t sublen;       /* L Booleans stored, m:26 (200)    */
    double   sample_rate;  /* m:128 (5512.0)                   */
} lbad_oracle_cfg;

#define LBAD_ORACLE_ROWS_PER_FRAME 128u   /* m:25, hard-coded in the reference */

void     lbad_oracle_default_cfg(lbad_oracle_cfg* c);
/* m:361-371, m:380-383: idx[bands+1], klow[bands], khigh[bands] for a window of nframes samples */
void     lbad_oracle_band_table(const lbad_oracle_cfg* c, uint32_t nframes, uint32_t* idx, uint32_t* klow, uint32_t* khigh);
/* m:353-355 as DEFINED for vDSP: out[2k],out[2k+1] = 2 Re X[k], 2 Im X[k]; out[0]=2X[0], out[1]=2X[N/2] */
void     lbad_oracle_fft2x(const float* x, uint32_t n, float* out);
/* m:335-408 on one window */
void     lbad_oracle_window_bands(const lbad_oracle_cfg* c, const float* win, float* out);
/* Frame.m:113-153 */
void     lbad_oracle_haar(float* image, uint32_t rows, uint32_t cols);
/* Frame.m:165-191 with the stable-sort definition; out: 2*t Booleans, zeroed here */
void     lbad_oracle_extract_bits(const float* coeffs, uint32_t n, uint32_t t, uint8_t* out);
/* m:250-331: number of subfingerprints for a clip of n samples (0 if n < window) */
uint64_t lbad_oracle_subfp_count(const lbad_oracle_cfg* c, int64_t n);
/* whole path; out_bits [count][sublen]; optional stage dumps [count][128][bands] */
int      lbad_oracle_process(const lbad_oracle_cfg* c, const float* pcm, int64_t n, uint8_t* out_bits, uint32_t max_subfps,
                             uint32_t* out_count, float* out_images, float* out_haar);
/* FP.m:151-176 ; len1 = subfingerprintLength of the fingerprint that owns s1 */
float    lbad_oracle_compare_sub(const uint8_t* s1, const uint8_t* s2, uint32_t len1, uint32_t range);
/* FP.m:119-149 ; bits [count][L] */
float    lbad_oracle_compare_fp(const uint8_t* b1, uint32_t c1, uint32_t l1, const uint8_t* b2, uint32_t c2, uint32_t l2, uint32_t range);
/* m:442-464 */
int      lbad_oracle_compare_pcm(const lbad_oracle_cfg* c, const float* p1, int64_t n1, const float* p2, int64_t n2, uint32_t range, float* out);

/* host-thread batch runners for the CPU baseline ("port" kind); return wall seconds */
double   lbad_oracle_extract_batch(const lbad_oracle_cfg* c, const float* pcm, uint32_t n_clips, int64_t clip_len, uint32_t threads,
                                   uint8_t* out_bits, uint32_t max_subfps_per_clip, uint32_t* out_counts, int fft_f32);
double   lbad_oracle_search(const uint8_t* db_bits, uint32_t n_db, uint32_t db_count, const uint8_t* q_bits, uint32_t n_q, uint32_t q_count,
                            uint32_t L, uint32_t range, uint32_t threads, float* scores);

/* recording-rate -> processing-rate conversion as DEFINED in include/LBAudioDetectiveResample.h (the reference leaves it to
 * ExtAudioFile, m:229 / m:275); out holds lbad_oracle_resampled_length() samples */
uint64_t lbad_oracle_resampled_length(double in_rate, double out_rate, uint64_t n_in);
uint64_t lbad_oracle_resample(double in_rate, double out_rate, const float* x, int64_t n_in, float* out);

/* deterministic synthetic PCM: chirp + tone + uniform noise, double arithmetic, f32 out */
void     lbad_synth_clip(uint64_t base_seed, uint64_t clip_id, int64_t n, double sample_rate, float* out);
void     lbad_synth_add_noise(uint64_t seed, int64_t n, double amplitude, float* io);

#ifdef __cplusplus
}
#endif
#endif
