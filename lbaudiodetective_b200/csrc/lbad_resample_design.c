/*
 * lbad_resample_design.c — filter design for the recording-rate -> processing-rate conversion (host, double precision,
 * rounded once to float32).  The definition is in include/LBAudioDetectiveResample.h; lbad_resample.cu evaluates the tables.
 */
#include "lbad_cuda.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static double bessel_i0(double x) {                      /* power series, converges fast for the beta values used (<= 10) */
    double sum = 1.0, term = 1.0, h = 0.5 * x;
    for (int k = 1; k < 200; k++) {
        term *= (h / k) * (h / k);
        sum += term;
        if (term < 1e-20 * sum) break;
    }
    return sum;
}
static double kaiser(double x, double beta) { return fabs(x) >= 1.0 ? 0.0 : bessel_i0(beta * sqrt(1.0 - x * x)) / bessel_i0(beta); }
static double sinc(double x) { return x == 0.0 ? 1.0 : sin(M_PI * x) / (M_PI * x); }

uint64_t lbad_resample_out_len(double in_rate, double out_rate, uint64_t n_in) {
    if (!(in_rate >= out_rate) || !(out_rate > 0.0)) return 0;
    return (uint64_t)floor((double)n_in * out_rate / in_rate + 1e-9);
}

int lbad_resample_design_create(double in_rate, double out_rate, lbadcu_resample_design* d) {
    memset(d, 0, sizeof *d);
    if (!(out_rate > 0.0) || !(in_rate >= out_rate) || in_rate / out_rate > 64.0) return LBAD_ERR_ARG;
    const double rho = in_rate / out_rate;
    uint32_t D = (uint32_t)floor(rho / 2.0);
    if (D < 1) D = 1;
    d->in_rate = in_rate; d->out_rate = out_rate; d->D = D; d->rho2 = rho / (double)D;
    if (D > 1) {
        d->H1 = 6 * D; d->T1 = 2 * d->H1 + 1;
        d->g = malloc(d->T1 * sizeof(float));
        double* t = malloc(d->T1 * sizeof(double));
        if (!d->g || !t) { free(t); lbad_resample_design_free(d); return LBAD_ERR_ARG; }
        double sum = 0.0;
        for (uint32_t i = 0; i < d->T1; i++) {
            const double u = (double)i - (double)d->H1;
            t[i] = sinc(u / (double)D) / (double)D * kaiser(u / (double)(d->H1 + 1), 8.0);
            sum += t[i];
        }
        for (uint32_t i = 0; i < d->T1; i++) d->g[i] = (float)(t[i] / sum);
        free(t);
    }
    const double gamma = 0.97 / d->rho2;
    d->H2 = (uint32_t)ceil(10.0 / gamma); d->T2 = 2 * d->H2;
    d->hc = malloc((size_t)(LBAD_RS_PHASES + 1) * d->T2 * sizeof(float));
    double* row = malloc(d->T2 * sizeof(double));
    if (!d->hc || !row) { free(row); lbad_resample_design_free(d); return LBAD_ERR_ARG; }
    for (uint32_t p = 0; p <= LBAD_RS_PHASES; p++) {
        double sum = 0.0;
        for (uint32_t i = 0; i < d->T2; i++) {
            const double u = (double)i - (double)d->H2 + 1.0 - (double)p / (double)LBAD_RS_PHASES;
            row[i] = fabs(u) < (double)d->H2 ? gamma * sinc(gamma * u) * kaiser(u / (double)d->H2, 9.0) : 0.0;
            sum += row[i];
        }
        for (uint32_t i = 0; i < d->T2; i++) d->hc[(size_t)p * d->T2 + i] = (float)(row[i] / sum);
    }
    free(row);
    return LBAD_OK;
}

void lbad_resample_design_free(lbadcu_resample_design* d) {
    if (!d) return;
    free(d->g); free(d->hc);
    d->g = NULL; d->hc = NULL;
}
