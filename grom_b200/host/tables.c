/* tables.c -- the two 1001x1001 binomial tail tables and the p-value -> sd table the
 * statistical scan looks up; host builds them once, the GPU library copies them.
 *
 * Restates reference src/GROM.c:21134-21586 (read_binom_tables), 21589-21626
 * (calculate_normal_binom_constants) and 20705-20748 (pval2sd list).  The
 * arithmetic keeps the reference's operation order and its integer types on
 * purpose: the factorial and the running binomial coefficient are C `long`s
 * that overflow / truncate in the reference (src/GROM.c:21237-21244,
 * 21275-21288) and the published tables contain the resulting values.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "gromhost.h"

#define NT 1000
#define ROW (NT + 1)

/* Abramowitz-Stegun 7.1.26 constants as typed in the reference (src/GROM.c:21157-21162) */
static const double AS_P = 0.3275911, AS_A1 = 0.254829592, AS_A2 = -0.284496736, AS_A3 = 1.421413741,
                    AS_A4 = -1.453152027, AS_A5 = 1.061405429;

/* P(X < successes) for X ~ Binomial(n, prob), by the reference's three regimes.
 * normal_min: smallest `successes` for which the normal approximation is taken
 * (17 for the hez table src/GROM.c:21247, 20 for the mq table src/GROM.c:21452). */
static double lower_cdf(long n, long successes, double prob, long normal_min)
{
    double cdf = 0;
    long k;
    if ((n >= 20 && prob <= 0.05) || (n >= 100 && n * prob <= 10)) {
        double lambda = n * prob;
        long k_factorial = 1;                      /* overflows for k > 20, as in the reference */
        for (k = 0; k < successes; k++) {
            if (k > 1) k_factorial = (long)((unsigned long)k_factorial * (unsigned long)k);
            cdf += pow(lambda, k) * exp(-lambda) / (double)k_factorial;
        }
    } else if (n * prob * (1 - prob) >= 5 && successes >= normal_min) {
        double sd = sqrt(n * prob * (1.0 - prob));
        double mean = n * prob;
        double z = (mean - successes + 0.5) / sd;
        double x = z / sqrt(2.0);
        double t = 1.0 / (1.0 + AS_P * x);
        double erf_ = 1.0 - (AS_A1 * t + AS_A2 * pow(t, 2) + AS_A3 * pow(t, 3) + AS_A4 * pow(t, 4) + AS_A5 * pow(t, 5)) * exp(-pow(x, 2));
        if (z >= 0) cdf = (1.0 - erf_) / 2.0;
        else        cdf = 1 - (erf_ + (1.0 - erf_) / 2.0);
    } else {
        long n_minus_k = n;
        long comb = 1;                             /* double expression truncated back to long each step */
        for (k = 0; k < successes; k++) {
            cdf += comb * pow(prob, k) * pow((1 - prob), n_minus_k);
            if (k > 0) comb = (long)((comb / (k + 1.0)) * n_minus_k);
            else       comb = comb * n_minus_k;
            n_minus_k -= 1;
        }
    }
    if (cdf < 0) cdf = 0;
    if (cdf > 1) cdf = 1;
    return cdf;
}

/* hez[n][k] (p = 0.5): rows 0..999 end up as P(X <= k); row 1000 stays P(X >= k) (src/GROM.c:21301-21316) */
static void compute_hez(double *hez)
{
    long n, s;
    int r, c;
    memset(hez, 0, sizeof(double) * ROW * ROW);
    for (n = 1; n <= NT; n++)
        for (s = 0; s <= n; s++)
            hez[n * ROW + s] = 1.0 - lower_cdf(n, s, 0.5, 17);
    for (r = 0; r < NT; r++) {
        double *row = hez + (size_t)r * ROW;
        for (c = 0; c < NT; c++) {
            row[c] = 1.0 - row[c + 1];
            if (row[c] < 0) row[c] = 0;
            if (c > 0 && row[c - 1] == 1) row[c] = 1;
        }
        row[NT] = 1.0;
    }
}

/* mq[n][k] (p = 10^(-q/10)) = P(X >= k), forced to 0 once the tail underflows or stalls (src/GROM.c:21431-21436) */
static void compute_mq(double *mq, double prob)
{
    long n, s;
    memset(mq, 0, sizeof(double) * ROW * ROW);
    for (n = 1; n <= NT; n++) {
        double *row = mq + (size_t)n * ROW;
        for (s = 0; s <= n; s++) {
            if ((s > 0 && row[s - 1] == 0) || (s > 1 && row[s - 1] == row[s - 2])) row[s] = 0;
            else row[s] = 1.0 - lower_cdf(n, s, prob, 20);
        }
    }
}

double gromhost_mq_prob(int min_mapq) { return pow(10, (-min_mapq / 10.0)); }   /* src/GROM.c:21612 */

void gromhost_tables_compute(int min_mapq, double *hez, double *mq)
{
    compute_hez(hez);
    compute_mq(mq, gromhost_mq_prob(min_mapq));
}

static int load_one(const char *path, double *tbl)
{
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    size_t cap = 100000;
    char *line = (char *)malloc(cap);
    int r = 0;
    memset(tbl, 0, sizeof(double) * ROW * ROW);
    while (r < ROW && fgets(line, (int)cap, f)) {
        char *save = NULL, *tok = strtok_r(line, "\t", &save);
        int c = 0;
        while (tok && c < ROW) { tbl[(size_t)r * ROW + c] = atof(tok); c++; tok = strtok_r(NULL, "\t", &save); }
        r++;
    }
    free(line); fclose(f);
    return r == ROW ? 0 : -2;
}

static int save_one(const char *path, const double *tbl)
{
    FILE *f = fopen(path, "w");
    if (!f) return -1;
    for (int r = 0; r < ROW; r++) {
        for (int c = 0; c < ROW; c++) { fprintf(f, "%e", tbl[(size_t)r * ROW + c]); if (c < NT) fputc('\t', f); }
        fputc('\n', f);
    }
    fclose(f);
    return 0;
}

void gromhost_table_paths(const char *dir, int min_mapq, char *hez_path, char *mq_path, int cap)
{
    snprintf(hez_path, cap, "%s/GROM_hez_binom_table_%d.txt", dir, NT);
    snprintf(mq_path, cap, "%s/GROM_mq_binom_table_%d_%d.txt", dir, min_mapq > 10 ? min_mapq : 10, NT);
}

/* Reference behaviour (src/GROM.c:21210-21375, 21400-21581): load each table from the text file next
 * to the executable when it exists (values then carry the 7 digits of "%e"), otherwise compute it and
 * try to write the file.  write_missing = 0 suppresses the write. */
int gromhost_tables_get(const char *dir, int min_mapq, int write_missing, double *hez, double *mq)
{
    char hp[4096], mp[4096];
    gromhost_table_paths(dir ? dir : ".", min_mapq, hp, mp, sizeof(hp));
    if (!dir || load_one(hp, hez) != 0) { compute_hez(hez); if (dir && write_missing) save_one(hp, hez); }
    if (!dir || load_one(mp, mq) != 0) { compute_mq(mq, gromhost_mq_prob(min_mapq)); if (dir && write_missing) save_one(mp, mq); }
    return 0;
}

/* p-value -> number of standard deviations, 1001 entries from sd = 10.00 down to 0.00 in steps of
 * g_stdev_step = 0.01 (src/GROM.c:20705-20748); pval[] ascending in index, sd[] descending. */
int gromhost_pval2sd(double *pval, double *sd, int cap)
{
    const double sd_max = 10.0, step = 0.01;
    int len = (int)(sd_max / step + 0.5) + 1;
    if (cap < len) return -1;
    for (int i = 0; i < len; i++) {
        double s = sd_max - i * step;
        if (s < 0) s = 0;
        double x = s / sqrt(2.0);
        double t = 1.0 / (1.0 + AS_P * x);
        double erf_ = 1.0 - ((AS_A1 * t + AS_A2 * pow(t, 2) + AS_A3 * pow(t, 3) + AS_A4 * pow(t, 4) + AS_A5 * pow(t, 5)) * exp(-pow(x, 2)));
        pval[i] = (1.0 - erf_) / 2.0;
        sd[i] = s;
    }
    return len;
}
