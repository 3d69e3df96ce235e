/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference LengthRegulator.
 *
 * Restates models/tts/fastspeech2/layers.py:434-462 (LengthRegulator.forward) and the
 * pad_list it calls (models/tts/fastspeech2/function.py:97-124, byte-identical to
 * espnet.nets.pytorch_backend.nets_utils.pad_list imported at layers.py:11).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load the shared
 * object built from this file (oracle/Makefile -> oracle/_build/liblr_oracle.so).
 * Pinned by tests/test_oracle_pinned.py against the unmodified reference and by
 * tests/golden/lr_*.npz.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* layers.py:446-448: ds = torch.round(ds.float() * alpha).long()  (round-half-to-even) */
void lr_oracle_scale(const int64_t *ds, int64_t n, float alpha, int64_t *out) {
    for (int64_t i = 0; i < n; ++i) out[i] = (int64_t)nearbyintf((float)ds[i] * alpha);
}

/* layers.py:450-458: whole-batch sum == 0 -> every element of every all-zero row becomes 1,
 * in place.  Returns 1 when the fix-up was applied. */
int lr_oracle_fix_all_zero(int64_t *ds, int64_t B, int64_t Tmax) {
    int64_t total = 0;
    for (int64_t i = 0; i < B * Tmax; ++i) total += ds[i];
    if (total != 0) return 0;
    for (int64_t b = 0; b < B; ++b) {
        int64_t s = 0;
        for (int64_t t = 0; t < Tmax; ++t) s += ds[b * Tmax + t];
        if (s == 0) for (int64_t t = 0; t < Tmax; ++t) ds[b * Tmax + t] = 1;
    }
    return 1;
}

/* layers.py:209 (caller): mel_lens = sum(ds, dim=1); returns max (pad_list max_len), or -1 on
 * a negative duration (torch.repeat_interleave rejects those). */
int64_t lr_oracle_rowsum(const int64_t *ds, int64_t B, int64_t Tmax, int64_t *mel_len) {
    int64_t mx = 0;
    for (int64_t b = 0; b < B; ++b) {
        int64_t s = 0;
        for (int64_t t = 0; t < Tmax; ++t) {
            if (ds[b * Tmax + t] < 0) return -1;
            s += ds[b * Tmax + t];
        }
        mel_len[b] = s;
        if (s > mx) mx = s;
    }
    return mx;
}

/* layers.py:460 + function.py:117-122: repeat_interleave each row, then pad to T_out.
 * Elements are moved as opaque `esize`-byte items; `pad` points at one element's bytes. */
void lr_oracle_expand(const void *xs, const int64_t *ds, void *out, int64_t B, int64_t Tmax,
                      int64_t D, int64_t T_out, int64_t esize, const void *pad) {
    const char *x = (const char *)xs;
    char *o = (char *)out;
    const int64_t row = D * esize;
    for (int64_t b = 0; b < B; ++b) {
        int64_t t = 0;
        for (int64_t i = 0; i < Tmax; ++i)
            for (int64_t r = 0; r < ds[b * Tmax + i] && t < T_out; ++r, ++t)
                memcpy(o + (b * T_out + t) * row, x + (b * Tmax + i) * row, (size_t)row);
        for (; t < T_out; ++t)
            for (int64_t d = 0; d < D; ++d)
                memcpy(o + (b * T_out + t) * row + d * esize, pad, (size_t)esize);
    }
}
