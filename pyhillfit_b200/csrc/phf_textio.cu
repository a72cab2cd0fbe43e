// Host-side text writer for chain / sample files: the bytes numpy.savetxt(fmt='%.18e') produces, written by all
// host cores.  The reference saves every chain with np.savetxt (python/PyHillFit.py:514-515, 866-867;
// python/PyHillTemp.py:169) at ~17 MB/s per process; a default thermodynamic-integration sweep is ~130 GB of text
// (SURVEY.md section 7), so after the kernels this is the end-to-end bottleneck.  Formatting uses the C library's
// correctly rounded "%.18e" (identical digits to Python's), one contiguous row range per thread, then one write.
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "phf_common.cuh"

using namespace phf;

namespace {

void format_range(const double *data, int64_t row_begin, int64_t row_end, int32_t n_cols, int64_t row_stride,
                  std::string *out)
{
    out->clear();
    out->reserve((size_t)(row_end - row_begin) * (size_t)n_cols * 26);
    char buf[64];
    for (int64_t r = row_begin; r < row_end; ++r) {
        const double *row = data + r * row_stride;
        for (int32_t c = 0; c < n_cols; ++c) {
            const double v = row[c];
            int len;
            if (std::isnan(v)) {  // Python prints "nan" for either sign
                memcpy(buf, "nan", 3);
                len = 3;
            } else {
                len = snprintf(buf, sizeof buf, "%.18e", v);
            }
            out->append(buf, (size_t)len);
            out->push_back(c + 1 < n_cols ? ' ' : '\n');
        }
    }
}

}  // namespace

// Write `header` (may be NULL; written verbatim, the caller includes '#' and newlines) followed by n_rows x n_cols
// numbers from a HOST array whose rows are `row_stride` doubles apart.  append != 0 appends to an existing file.
// n_threads <= 0: all hardware threads.  Returns PHF_OK or PHF_EINVAL (text via phf_last_error).
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
    std::vector<std::string> parts((size_t)nt);
    std::vector<std::thread> pool;
    for (int k = 0; k < nt; ++k) {
        const int64_t b = n_rows * k / nt, e = n_rows * (k + 1) / nt;
        if (k + 1 < nt)
            pool.emplace_back(format_range, data, b, e, n_cols, row_stride, &parts[(size_t)k]);
        else
            format_range(data, b, e, n_cols, row_stride, &parts[(size_t)k]);
    }
    for (auto &t : pool) t.join();
    FILE *f = fopen(path, append ? "ab" : "wb");
    if (!f) {
        char msg[400];
        snprintf(msg, sizeof msg, "phf_write_rows_text_host: cannot open %.300s: %s", path, strerror(errno));
        return set_error(PHF_EINVAL, msg);
    }
    bool ok = true;
    if (header && *header) ok = fwrite(header, 1, strlen(header), f) == strlen(header);
    for (int k = 0; k < nt && ok; ++k)
        ok = fwrite(parts[(size_t)k].data(), 1, parts[(size_t)k].size(), f) == parts[(size_t)k].size();
    ok = (fclose(f) == 0) && ok;
    if (!ok) return set_error(PHF_EINVAL, "phf_write_rows_text_host: short write");
    return PHF_OK;
}
