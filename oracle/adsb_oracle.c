/*
 * CPU oracle, C restatement -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Restates gr-adsb python/adsb/framer.py:72-182 and python/adsb/demod.py:57-136 in plain C so
 * that parity tests can check the CUDA path on streams of 10^7..10^8 samples in seconds.  It
 * is pinned the same way as oracle/adsb_oracle.py: against the golden vectors under
 * tests/golden/ (outputs of the unmodified reference) and, where /root/reference exists,
 * against the live reference.  Build: oracle/Makefile (gcc -O2 -ffp-contract=off).
 *
 * Float semantics follow NumPy >= 2 on float32 input: every comparison, the x[p]/2 half
 * level, the median's (a+b)/2 and the x[p]/median ratio are float32 operations
 * (framer.py:84,141,157-159); log10 is libm's log10f (NumPy may use a SIMD variant that
 * differs in the last ulp -- covered by the 1e-4 dB SNR tolerance of SURVEY.md 8c).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PREAMBLE_SYMBOLS 8   /* framer.py:28 */
#define SHORT_MSG_BITS 56    /* framer.py:29 */
#define NOISE_WINDOW 100     /* framer.py:31 */
#define MSG_BITS 112         /* demod.py:29  */

typedef struct {
    float prev_in0;      /* framer.py:54 */
    int64_t prev_eob;    /* framer.py:57 */
} adsb_oracle_state;

static const int TEMPLATE[16] = {1, 0, 1, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 0, 0}; /* framer.py:50 */

void adsb_oracle_state_init(adsb_oracle_state *st)
{
    st->prev_in0 = 0.0f;
    st->prev_eob = -1;
}

static int cmp_float(const void *a, const void *b)
{
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

/* np.median of w[0..n): nan when empty or any nan; even n -> fl32(fl32(a+b)/2) */
static float median_f32(const float *w, int n)
{
    float tmp[NOISE_WINDOW];
    if (n <= 0)
        return NAN;
    for (int i = 0; i < n; i++) {
        if (isnan(w[i]))
            return NAN;
        tmp[i] = w[i];
    }
    qsort(tmp, (size_t)n, sizeof(float), cmp_float);
    if (n & 1)
        return tmp[n / 2];
    volatile float s = tmp[n / 2 - 1] + tmp[n / 2];
    volatile float m = s / 2.0f;
    return m;
}

/* One framer.work() call.  in0 has n_out + 8*sps - 1 items.  Returns the number of hits
 * (window coordinates p ascending, SNR in dB), or -1 if more than `cap` were found. */
int64_t adsb_oracle_framer_work(adsb_oracle_state *st, const float *in0, int64_t n_out, int sps,
                                float threshold, int64_t *hit_p, float *hit_snr, int64_t cap)
{
    const int64_t N = n_out;
    const int half = sps / 2;
    const int before = (st->prev_in0 >= threshold);   /* bitmap value at i = -1 (framer.py:83-84) */
    int64_t n_hits = 0, n_rise = 0, n_fall = 0;

    /* framer.py:91-95: the call does nothing unless it holds at least one rise and one fall */
    int prev = before;
    for (int64_t i = 0; i < N; i++) {
        int cur = (in0[i] >= threshold);
        n_rise += (cur && !prev);
        n_fall += (!cur && prev);
        prev = cur;
    }
    st->prev_in0 = in0[N - 1];                                   /* framer.py:87 */
    if (n_rise == 0 || n_fall == 0)                              /* framer.py:95 */
        return 0;

    /* framer.py:98-113: complete pulses [rise, fall) inside [0, N); a run already high at i = -1
     * has no rise here (its fall is the discarded leading one), a run still high at N-1 has no fall */
    int64_t rise = -1;
    prev = before;
    for (int64_t i = 0; i < N; i++) {
        int cur = (in0[i] >= threshold);
        if (cur && !prev)
            rise = i;
        if (!cur && prev) {
            if (rise >= 0) {
                int64_t p = (rise + i) / 2;                      /* framer.py:113 */
                if (p > st->prev_eob) {                          /* framer.py:121 */
                    st->prev_eob = -1;                           /* framer.py:123 */
                    float level = in0[p] / 2.0f;                 /* framer.py:141 */
                    int ok = 1;
                    for (int k = 0; k < 16 && ok; k++)           /* framer.py:137-147 */
                        ok = ((in0[p + (int64_t)k * half] > level) == TEMPLATE[k]);
                    if (ok) {
                        int64_t lo = p < NOISE_WINDOW ? 0 : p - NOISE_WINDOW;   /* framer.py:156-159 */
                        float med = median_f32(in0 + lo, (int)(p - lo));
                        volatile float ratio = in0[p] / med;
                        volatile float lg = log10f(ratio);
                        volatile float s10 = 10.0f * lg;
                        volatile float snr = s10 + 1.6f;
                        st->prev_eob = p + (int64_t)(PREAMBLE_SYMBOLS + SHORT_MSG_BITS - 1) * sps; /* :165 */
                        if (n_hits >= cap)
                            return -1;
                        hit_p[n_hits] = p;
                        hit_snr[n_hits] = snr;
                        n_hits++;
                    }
                }
            }
            rise = -1;
        }
        prev = cur;
    }
    if (st->prev_eob >= N)                                       /* framer.py:177-179 */
        st->prev_eob -= N;
    return n_hits;
}

/* demod.py:87-95: 112 single-sample decisions, bit = (in0[sob + k*sps] > in0[sob + sps/2 + k*sps]) */
void adsb_oracle_demod_slice(const float *in0, int64_t sob, int sps, uint8_t *bits)
{
    const int half = sps / 2;
    for (int k = 0; k < MSG_BITS; k++)
        bits[k] = in0[sob + (int64_t)k * sps] > in0[sob + half + (int64_t)k * sps];
}

/* Lock-step framer -> demod over work() calls of `chunk` items (chunk <= 0: one call), history of
 * 8*sps-1 zeros at stream start.  Outputs: every tag (absolute offset, snr);
 * every PDU (offset, snr, 112 bits one per byte).  Returns 0, or -1 on capacity overflow. */
int adsb_oracle_run_stream(const float *x, int64_t n, int sps, float threshold, int64_t chunk,
                           int64_t *tag_off, float *tag_snr, int64_t *n_tags,
                           int64_t *pdu_off, float *pdu_snr, uint8_t *pdu_bits, int64_t *n_pdus,
                           int64_t cap, float *final_prev_in0, int64_t *final_prev_eob)
{
    const int64_t H = (int64_t)PREAMBLE_SYMBOLS * sps;
    adsb_oracle_state st;
    adsb_oracle_state_init(&st);
    if (chunk <= 0 || chunk > n)
        chunk = n;
    float *win = (float *)malloc(sizeof(float) * (size_t)(chunk + H - 1));
    int64_t *hp = (int64_t *)malloc(sizeof(int64_t) * (size_t)(chunk / 2 + 2));
    float *hs = (float *)malloc(sizeof(float) * (size_t)(chunk / 2 + 2));
    int64_t nt = 0, np_ = 0, pos = 0;
    int rc = 0;
    memset(win, 0, sizeof(float) * (size_t)(H - 1));
    while (pos < n && rc == 0) {
        int64_t m = n - pos < chunk ? n - pos : chunk;
        if (pos > 0) {
            /* history = last H-1 items before pos (zeros where the stream has not started) */
            for (int64_t j = 0; j < H - 1; j++) {
                int64_t src = pos - (H - 1) + j;
                win[j] = src >= 0 ? x[src] : 0.0f;
            }
        }
        memcpy(win + (H - 1), x + pos, sizeof(float) * (size_t)m);
        int64_t nh = adsb_oracle_framer_work(&st, win, m, sps, threshold, hp, hs, chunk / 2 + 2);
        for (int64_t i = 0; i < nh; i++) {
            int64_t off = pos - (H - 1) + hp[i];
            if (nt >= cap) { rc = -1; break; }
            tag_off[nt] = off;
            tag_snr[nt] = hs[i];
            nt++;
            /* demod.py:67,82: tag must lie inside this buffer and the burst must end inside it */
            int64_t q = off - pos;
            if (q >= 0 && 2 * (q + (int64_t)(PREAMBLE_SYMBOLS + MSG_BITS - 1) * sps) + sps < 2 * m) {
                if (np_ >= cap) { rc = -1; break; }
                pdu_off[np_] = off;
                pdu_snr[np_] = hs[i];
                adsb_oracle_demod_slice(x + pos, q + (int64_t)PREAMBLE_SYMBOLS * sps, sps,
                                        pdu_bits + np_ * MSG_BITS);
                np_++;
            }
        }
        pos += m;
    }
    *n_tags = nt;
    *n_pdus = np_;
    if (final_prev_in0) *final_prev_in0 = st.prev_in0;
    if (final_prev_eob) *final_prev_eob = st.prev_eob;
    free(win); free(hp); free(hs);
    return rc;
}

/* complex64 -> |IQ|^2, multiply then add in float32 (gr-blocks complex_to_mag_squared, adsb_rx.py:180) */
void adsb_oracle_mag2(const float *iq, int64_t n, float *out)
{
    for (int64_t i = 0; i < n; i++) {
        volatile float a = iq[2 * i] * iq[2 * i];
        volatile float b = iq[2 * i + 1] * iq[2 * i + 1];
        out[i] = a + b;
    }
}

/* Remainder of the 112-bit message polynomial modulo the Mode S generator 0x1FFF409
 * (decoder.py:269,693-715); 0 <=> PI == CRC(bits[0:88]) (decoder.py:673-680). */
uint32_t adsb_oracle_crc24_syndrome(const uint8_t *bits, int nbits)
{
    uint32_t r = 0;
    for (int i = 0; i < nbits; i++) {
        r = (r << 1) | (bits[i] & 1u);
        if (r & (1u << 24))
            r ^= 0x1FFF409u;
    }
    return r & 0xFFFFFFu;
}

/* decoder.check_parity() under the "Extended Squitter Only" filter (decoder.py:667-685) */
int adsb_oracle_crc_pass(const uint8_t *bits112)
{
    int df = (bits112[0] << 4) | (bits112[1] << 3) | (bits112[2] << 2) | (bits112[3] << 1) | bits112[4];
    if (df < 17 || df > 19)
        return 0;
    return adsb_oracle_crc24_syndrome(bits112, MSG_BITS) == 0;
}
