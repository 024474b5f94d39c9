/* oracle/mt19937.h - the reference's random stream, restated: MT19937 plus the GNU libstdc++ (GCC 13) distributions the
 * reference algorithms draw through.  TEST INFRASTRUCTURE ONLY.
 *
 * pagmo's detail::random_engine_type is std::mt19937 (reference include/pagmo/rng.hpp:51); every UDA owns one (`m_e`) and
 * draws through std::uniform_real_distribution<double>, std::uniform_int_distribution, std::normal_distribution,
 * std::binomial_distribution and std::shuffle.  The C++ standard fixes MT19937's output (10000th value of a default-seeded
 * engine = 4123659995) but NOT the distributions; the reference is built with GCC here, so the sequences below are the
 * published algorithms of libstdc++ 13 (bits/random.tcc, bits/uniform_int_dist.h, bits/stl_algo.h):
 *   u01        generate_canonical<double,53>: (w0 + w1 * 2^32) / 2^64, nextafter(1,0) if that rounds to 1
 *   uint       Lemire's nearly-divisionless reduction of one 32-bit word (ranges below 2^32 - 1)
 *   shuffle    Fisher-Yates from the front, two swap positions from one draw while range^2 fits in 32 bits
 *   normal     Marsaglia polar method, the spare deviate saved in the distribution object
 *   binomial   waiting-time algorithm for t*p < 8 (the only branch the reference's sga reaches at its defaults)
 * With a stream in this mode (oracle_stream.mt != NULL, philox.h) the restated operators consume exactly the draws the
 * compiled reference consumes, which is what tests/test_oracle_pin.py checks bit for bit against oracle/_ref; the
 * helpers themselves are pinned against the real std:: classes through ref_std_* (ref_capi.h).
 */
#ifndef ORACLE_MT19937_H
#define ORACLE_MT19937_H
#include <math.h>
#include <stddef.h>
#include <stdint.h>

typedef struct oracle_mt {
    uint32_t s[624];
    int pos;
    /* state of ONE std::normal_distribution object (the reference algorithms that draw normals own exactly one) */
    int saved_available;
    double saved;
} oracle_mt;

static inline void oracle_mt_seed(oracle_mt *m, uint32_t seed)
{
    m->s[0] = seed;
    for (int i = 1; i < 624; ++i) m->s[i] = 1812433253u * (m->s[i - 1] ^ (m->s[i - 1] >> 30)) + (uint32_t)i;
    m->pos = 624;
    m->saved_available = 0;
    m->saved = 0.;
}

static inline uint32_t oracle_mt_u32(oracle_mt *m)
{
    if (m->pos >= 624) {
        for (int k = 0; k < 624; ++k) {
            const uint32_t y = (m->s[k] & 0x80000000u) | (m->s[(k + 1) % 624] & 0x7fffffffu);
            m->s[k] = m->s[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        m->pos = 0;
    }
    uint32_t z = m->s[m->pos++];
    z ^= z >> 11;
    z ^= (z << 7) & 0x9d2c5680u;
    z ^= (z << 15) & 0xefc60000u;
    z ^= z >> 18;
    return z;
}

/* std::uniform_real_distribution<double>(0,1)(e) == generate_canonical<double,53>(e): two words, low word first */
static inline double oracle_mt_u01(oracle_mt *m)
{
    const double lo = (double)oracle_mt_u32(m);
    const double hi = (double)oracle_mt_u32(m);
    double r = (lo + hi * 4294967296.0) / 18446744073709551616.0;
    if (r >= 1.0) r = nextafter(1.0, 0.0);
    return r;
}

/* std::uniform_real_distribution<double>(a,b)(e) */
static inline double oracle_mt_real(oracle_mt *m, double a, double b) { return oracle_mt_u01(m) * (b - a) + a; }

/* std::uniform_int_distribution<size_t>(0, n-1)(e), n <= 2^32 - 1 */
static inline uint64_t oracle_mt_below(oracle_mt *m, uint64_t n)
{
    const uint32_t range = (uint32_t)n;
    uint64_t product = (uint64_t)oracle_mt_u32(m) * (uint64_t)range;
    uint32_t low = (uint32_t)product;
    if (low < range) {
        const uint32_t threshold = (uint32_t)(-range) % range;
        while (low < threshold) {
            product = (uint64_t)oracle_mt_u32(m) * (uint64_t)range;
            low = (uint32_t)product;
        }
    }
    return product >> 32;
}

/* std::uniform_int_distribution<T>(a, b)(e) for 0 <= b - a < 2^32 - 1 */
static inline uint64_t oracle_mt_int(oracle_mt *m, uint64_t a, uint64_t b) { return a + oracle_mt_below(m, b - a + 1u); }

/* std::shuffle(v, v + n, e) on an index array */
static inline void oracle_mt_shuffle(oracle_mt *m, size_t *v, size_t n)
{
    if (n == 0) return;
    const uint64_t urngrange = 0xffffffffull;
    size_t i = 1, t;
    if (urngrange / n >= n) {
        if ((n % 2) == 0) {
            const size_t j = (size_t)oracle_mt_int(m, 0, 1);
            t = v[i]; v[i] = v[j]; v[j] = t;
            ++i;
        }
        while (i != n) {
            const uint64_t swap_range = (uint64_t)i + 1u, b1 = swap_range + 1u;
            const uint64_t x = oracle_mt_int(m, 0, swap_range * b1 - 1u);
            const size_t j0 = (size_t)(x / b1), j1 = (size_t)(x % b1);
            t = v[i]; v[i] = v[j0]; v[j0] = t;
            ++i;
            t = v[i]; v[i] = v[j1]; v[j1] = t;
            ++i;
        }
        return;
    }
    for (; i != n; ++i) {
        const size_t j = (size_t)oracle_mt_int(m, 0, i);
        t = v[i]; v[i] = v[j]; v[j] = t;
    }
}

/* std::normal_distribution<double>(mean, stddev)(e) of the one distribution object whose spare lives in *m */
static inline double oracle_mt_normal(oracle_mt *m, double mean, double stddev)
{
    double ret;
    if (m->saved_available) {
        m->saved_available = 0;
        ret = m->saved;
    } else {
        double x, y, r2;
        do {
            x = 2.0 * oracle_mt_u01(m) - 1.0;
            y = 2.0 * oracle_mt_u01(m) - 1.0;
            r2 = x * x + y * y;
        } while (r2 > 1.0 || r2 == 0.0);
        const double mult = sqrt(-2 * log(r2) / r2);
        m->saved = x * mult;
        m->saved_available = 1;
        ret = y * mult;
    }
    return ret * stddev + mean;
}

/* std::binomial_distribution<size_t>(t, p)(e): the waiting-time branch (t * min(p, 1-p) < 8); returns (uint64_t)-1 otherwise */
static inline uint64_t oracle_mt_binomial(oracle_mt *m, uint64_t t, double p)
{
    const double p12 = p <= 0.5 ? p : 1.0 - p;
    if ((double)t * p12 >= 8) return (uint64_t)-1;
    const double q = -log(1 - p12);
    uint64_t x = 0, ret;
    double sum = 0.0;
    for (;;) {
        if (t == x) { ret = x; break; }
        const double e = -log(1.0 - oracle_mt_u01(m));
        sum += e / (double)(t - x);
        x += 1;
        if (!(sum <= q)) { ret = x - 1; break; }
    }
    if (p12 != p) ret = t - ret;
    return ret;
}

#endif
