// Host-side text writer for chain / sample files: the bytes numpy.savetxt(fmt='%.18e') produces, written by all
// host cores.  The reference saves every chain with np.savetxt (python/PyHillFit.py:514-515, 866-867;
// python/PyHillTemp.py:169) at ~17 MB/s per process; a default thermodynamic-integration sweep is ~130 GB of text
// (SURVEY.md section 7), so after the kernels this is the end-to-end bottleneck.
//
// "%.18e" is 19 significant digits, correctly rounded (ties to even on the exact value) -- what both glibc and
// Python print.  format_e18 produces exactly those bytes with integer arithmetic for 1e-9 <= |v| < 1e19 (every
// parameter and log-target a chain holds): v = m 2^e, digits = round(m 5^p 2^(e+p)) with p = 18 - floor(log10 |v|);
// m 5^p fits 128 bits for p <= 27, the shift and its remainder are exact, so the rounding is exact.  Everything
// else (0, subnormals, tiny, huge, inf, nan) goes through the C library.  tests/test_host_logic.py checks it
// against the C library on 2e7 doubles (phf_format_e18_mismatches) and against Python's own formatting.
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <fcntl.h>
#include <unistd.h>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "phf_common.cuh"

using namespace phf;

namespace {

typedef unsigned __int128 u128;

const uint64_t kPow5[28] = {1ull, 5ull, 25ull, 125ull, 625ull, 3125ull, 15625ull, 78125ull, 390625ull, 1953125ull,
                            9765625ull, 48828125ull, 244140625ull, 1220703125ull, 6103515625ull, 30517578125ull,
                            152587890625ull, 762939453125ull, 3814697265625ull, 19073486328125ull,
                            95367431640625ull, 476837158203125ull, 2384185791015625ull, 11920928955078125ull,
                            59604644775390625ull, 298023223876953125ull, 1490116119384765625ull,
                            7450580596923828125ull};
const uint64_t kTen18 = 1000000000000000000ull, kTen19 = 10000000000000000000ull;

const char kDigitPairs[201] =
    "00010203040506070809101112131415161718192021222324252627282930313233343536373839"
    "40414243444546474849505152535455565758596061626364656667686970717273747576777879"
    "8081828384858687888990919293949596979899";

inline void put8(uint32_t v, char *p)  // 8 decimal digits
{
    const uint32_t a = v / 10000u, b = v % 10000u;
    memcpy(p, kDigitPairs + 2 * (a / 100u), 2);
    memcpy(p + 2, kDigitPairs + 2 * (a % 100u), 2);
    memcpy(p + 4, kDigitPairs + 2 * (b / 100u), 2);
    memcpy(p + 6, kDigitPairs + 2 * (b % 100u), 2);
}

// round(m 5^p 2^sh) with ties to even; false if it does not fit the fast path
inline bool scaled_digits(uint64_t m, int e2, int p, uint64_t *digits)
{
    if (p < 0 || p > 27) return false;
    const u128 n = (u128)m * kPow5[p];
    const int sh = e2 + p;  // value = n 2^sh
    if (sh >= 0) {
        if (sh > 12) return false;  // m < 2^53, 5^p >= 1: n 2^sh < 1e19 needs sh <= 10 anyway
        const u128 d = n << sh;
        if (d >= (u128)kTen19 * 2) return false;
        *digits = (uint64_t)d;
        return d < ((u128)1 << 64);
    }
    const int s = -sh;
    if (s >= 127) return false;
    u128 d = n >> s;
    const u128 rem = n & (((u128)1 << s) - 1), half = (u128)1 << (s - 1);
    if (rem > half || (rem == half && (d & 1))) ++d;
    if (d >= ((u128)1 << 64)) return false;
    *digits = (uint64_t)d;
    return true;
}

// "%.18e" of v into buf (at least 32 bytes); returns the length
int format_e18(double v, char *buf)
{
    uint64_t bits;
    memcpy(&bits, &v, 8);
    const int bexp = (int)((bits >> 52) & 0x7ff);
    if (bexp == 0x7ff && (bits & 0xfffffffffffffull)) {  // Python prints "nan" for either sign
        memcpy(buf, "nan", 3);
        return 3;
    }
    if (bexp != 0 && bexp != 0x7ff) {
        const uint64_t m = (bits & 0xfffffffffffffull) | (1ull << 52);
        const int e2 = bexp - 1075;  // |v| = m 2^e2, 2^52 <= m < 2^53
        // floor(log10 |v|) from the binary exponent: |v| in [2^(e2+52), 2^(e2+53)); 78913 / 2^18 = log10(2) (+1e-6)
        int k = (int)(((int64_t)(e2 + 52) * 78913) >> 18);
        uint64_t d = 0;
        bool ok = false;
        for (int attempt = 0; attempt < 3; ++attempt) {
            ok = scaled_digits(m, e2, 18 - k, &d);
            if (!ok) break;
            if (d >= kTen19) { ++k; ok = false; continue; }  // k too small, or the rounding carried into a 20th digit
            if (d < kTen18) { --k; ok = false; continue; }
            break;
        }
        if (ok) {
            char *p = buf;
            if (bits >> 63) *p++ = '-';
            const uint64_t top = d / 10000000000000000ull, low = d % 10000000000000000ull;  // 3 + 16 digits
            const uint32_t t = (uint32_t)top;
            *p++ = (char)('0' + t / 100u);
            *p++ = '.';
            memcpy(p, kDigitPairs + 2 * (t % 100u), 2);
            p += 2;
            put8((uint32_t)(low / 100000000ull), p);
            put8((uint32_t)(low % 100000000ull), p + 8);
            p += 16;
            *p++ = 'e';
            int ke = k;
            *p++ = ke < 0 ? '-' : '+';
            if (ke < 0) ke = -ke;
            if (ke >= 100) {
                *p++ = (char)('0' + ke / 100);
                ke %= 100;
            }
            memcpy(p, kDigitPairs + 2 * ke, 2);
            p += 2;
            return (int)(p - buf);
        }
    }
    return snprintf(buf, 32, "%.18e", v);
}

// pwrite the whole buffer at `offset`
bool write_all(int fd, const char *p, size_t n, off_t offset)
{
    while (n > 0) {
        const ssize_t w = pwrite(fd, p, n, offset);
        if (w < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        p += w;
        n -= (size_t)w;
        offset += w;
    }
    return true;
}

struct TextPart {
    std::unique_ptr<char[]> buf;  // not value-initialised: the formatter writes every byte it reports
    size_t size = 0;
};

void format_range(const double *data, int64_t row_begin, int64_t row_end, int32_t n_cols, int64_t row_stride,
                  TextPart *out)
{
    // at most 26 bytes per number ("-d.(18 digits)e+ddd") plus its separator
    out->buf.reset(new char[(size_t)(row_end - row_begin) * (size_t)n_cols * 27 + 32]);
    char *p = out->buf.get();
    for (int64_t r = row_begin; r < row_end; ++r) {
        const double *row = data + r * row_stride;
        for (int32_t c = 0; c < n_cols; ++c) {
            p += format_e18(row[c], p);
            *p++ = c + 1 < n_cols ? ' ' : '\n';
        }
    }
    out->size = (size_t)(p - out->buf.get());
}

}  // namespace

// Write `header` (may be NULL; written verbatim, the caller includes '#' and newlines) followed by n_rows x n_cols
// numbers from a HOST array whose rows are `row_stride` doubles apart.  append != 0 appends to an existing file.
// n_threads <= 0: all hardware threads.  Every thread formats one contiguous row range and then writes it at its own
// file offset (pwrite), so neither the formatting nor the copy into the page cache is serial.
// Returns PHF_OK or PHF_EINVAL (text via phf_last_error).
extern "C" int phf_write_rows_text_host(const char *path, const char *header, const double *data, int64_t n_rows,
                                        int32_t n_cols, int64_t row_stride, int32_t append, int32_t n_threads)
{
    if (!path || n_rows < 0 || n_cols <= 0 || row_stride < n_cols || (n_rows > 0 && !data))
        return set_error(PHF_EINVAL, "phf_write_rows_text_host: bad argument");
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    const int64_t min_rows_per_thread = 2048;
    if ((int64_t)nt > (n_rows + min_rows_per_thread - 1) / min_rows_per_thread)
        nt = (int)((n_rows + min_rows_per_thread - 1) / min_rows_per_thread);
    if (nt < 1) nt = 1;
    std::vector<TextPart> parts((size_t)nt);
    {
        std::vector<std::thread> pool;
        for (int k = 0; k < nt; ++k) {
            const int64_t b = n_rows * k / nt, e = n_rows * (k + 1) / nt;
            if (k + 1 < nt)
                pool.emplace_back(format_range, data, b, e, n_cols, row_stride, &parts[(size_t)k]);
            else
                format_range(data, b, e, n_cols, row_stride, &parts[(size_t)k]);
        }
        for (auto &t : pool) t.join();
    }
    const int fd = open(path, O_WRONLY | O_CREAT | (append ? 0 : O_TRUNC), 0666);
    if (fd < 0) {
        char msg[400];
        snprintf(msg, sizeof msg, "phf_write_rows_text_host: cannot open %.300s: %s", path, strerror(errno));
        return set_error(PHF_EINVAL, msg);
    }
    off_t base = 0;
    if (append) {
        base = lseek(fd, 0, SEEK_END);
        if (base < 0) base = 0;
    }
    bool ok = true;
    if (header && *header) {
        ok = write_all(fd, header, strlen(header), base);
        base += (off_t)strlen(header);
    }
    std::vector<off_t> offset((size_t)nt);
    for (int k = 0; k < nt; ++k) {
        offset[(size_t)k] = base;
        base += (off_t)parts[(size_t)k].size;
    }
    if (ok) {
        std::vector<char> done((size_t)nt, 0);
        std::vector<std::thread> pool;
        auto job = [&](int k) { done[(size_t)k] = write_all(fd, parts[(size_t)k].buf.get(), parts[(size_t)k].size, offset[(size_t)k]); };
        for (int k = 0; k + 1 < nt; ++k) pool.emplace_back(job, k);
        job(nt - 1);
        for (auto &t : pool) t.join();
        for (int k = 0; k < nt; ++k) ok = ok && done[(size_t)k];
    }
    ok = (close(fd) == 0) && ok;
    if (!ok) return set_error(PHF_EINVAL, "phf_write_rows_text_host: short write");
    return PHF_OK;
}

// Test hooks: "%.18e" of one double through the writer's formatter (buf: at least 32 bytes; returns the length), and
// the number of values among data[0..n) whose formatted bytes differ from the C library's snprintf("%.18e").
extern "C" int phf_format_e18(double v, char *buf) { return format_e18(v, buf); }

extern "C" int64_t phf_format_e18_mismatches(const double *data, int64_t n)
{
    int64_t bad = 0;
    char a[64], b[64];
    for (int64_t i = 0; i < n; ++i) {
        const int la = format_e18(data[i], a);
        int lb;
        if (std::isnan(data[i])) {
            memcpy(b, "nan", 3);
            lb = 3;
        } else {
            lb = snprintf(b, sizeof b, "%.18e", data[i]);
        }
        if (la != lb || memcmp(a, b, (size_t)la) != 0) ++bad;
    }
    return bad;
}
