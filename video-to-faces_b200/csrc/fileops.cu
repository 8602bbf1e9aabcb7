// Output stage of the grouping path: the reference copies, moves and deletes one file per
// Python loop iteration (grouping.py:78-84 os.replace, :122-137 shutil.copyfile + os.remove, dupes.py:70-79).  Once the
// arithmetic takes seconds these loops are the wall-clock floor: ~70 us per copyfile in CPython (open/fstat/sendfile/close
// behind several layers of Python), i.e. more than a minute per million faces.  Host code only (no kernels): a batch of
// independent file operations run by native threads -- no GIL, no per-file interpreter work; the copy itself is
// copy_file_range(2) (in-kernel, reflink / server-side copy where the file system has one) with a read/write fallback.
#include <atomic>
#include <cerrno>
#include <cstring>
#include <fcntl.h>
#include <stdio.h>
#include <sys/sendfile.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <thread>
#include <unistd.h>
#include <vector>
#include "common.cuh"

namespace vtf {

static int copy_one(const char* src, const char* dst) {
    const int in = open(src, O_RDONLY | O_CLOEXEC);
    if (in < 0) return errno;
    struct stat st;
    if (fstat(in, &st) != 0) { const int e = errno; close(in); return e; }
    // shutil.copyfile: contents only, destination created with the default mode and truncated
    const int out = open(dst, O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0666);
    if (out < 0) { const int e = errno; close(in); return e; }
    int err = 0;
    off_t left = st.st_size;
    bool use_cfr = true;
    while (left > 0 && !err) {
        if (use_cfr) {
            const ssize_t n = sendfile(out, in, nullptr, (size_t)left);     // what shutil.copyfile itself uses on Linux
            if (n > 0) { left -= n; continue; }
            if (n == 0) break;                                          // source shrank
            if (errno == EXDEV || errno == ENOSYS || errno == EINVAL || errno == EOPNOTSUPP || errno == EPERM) { use_cfr = false; continue; }
            if (errno == EINTR) continue;
            err = errno;
        } else {
            char buf[1 << 16];
            const ssize_t n = read(in, buf, sizeof(buf));
            if (n == 0) break;
            if (n < 0) { if (errno == EINTR) continue; err = errno; break; }
            ssize_t w = 0;
            while (w < n) {
                const ssize_t m = write(out, buf + w, (size_t)(n - w));
                if (m < 0) { if (errno == EINTR) continue; err = errno; break; }
                w += m;
            }
            left -= n;
        }
    }
    if (st.st_size == 0 && !err) {                                      // /proc-like sources report size 0: stream until EOF
        char buf[1 << 16];
        for (;;) {
            const ssize_t n = read(in, buf, sizeof(buf));
            if (n == 0) break;
            if (n < 0) { if (errno == EINTR) continue; err = errno; break; }
            if (write(out, buf, (size_t)n) != n) { err = errno ? errno : EIO; break; }
        }
    }
    close(in);
    if (close(out) != 0 && !err) err = errno;
    return err;
}

}  // namespace vtf

extern "C" {

// op: 0 = copy src -> dst (shutil.copyfile), 1 = move src -> dst (os.replace), 2 = remove src (os.remove).
// Paths are NUL-terminated strings inside a blob, path i at blob + off[i].  Every operation is attempted; returns the
// number of failures, with the lowest failing index and its errno in *first_index / *first_errno.
int64_t vtf_file_ops(int op, const char* src_blob, const int64_t* src_off, const char* dst_blob, const int64_t* dst_off,
                     int64_t n, int threads, int64_t* first_index, int* first_errno) {
    if (op < 0 || op > 2 || !src_blob || !src_off || (op != 2 && (!dst_blob || !dst_off)) || n < 0) return -1;
    if (first_index) *first_index = -1;
    if (first_errno) *first_errno = 0;
    if (n == 0) return 0;
    if (threads < 1) threads = 1;
    if ((int64_t)threads > n) threads = (int)n;
    std::atomic<int64_t> next{0}, failed{0}, first{INT64_MAX};
    std::vector<int> err_of_first(threads, 0);
    std::vector<int64_t> idx_of_first(threads, INT64_MAX);
    const int64_t chunk = 64;
    auto work = [&](int tid) {
        for (;;) {
            const int64_t b = next.fetch_add(chunk);
            if (b >= n) break;
            const int64_t e = b + chunk < n ? b + chunk : n;
            for (int64_t i = b; i < e; ++i) {
                const char* s = src_blob + src_off[i];
                int err = 0;
                if (op == 0) err = vtf::copy_one(s, dst_blob + dst_off[i]);
                else if (op == 1) err = rename(s, dst_blob + dst_off[i]) == 0 ? 0 : errno;
                else err = unlink(s) == 0 ? 0 : errno;
                if (err) {
                    failed.fetch_add(1);
                    if (i < idx_of_first[tid]) { idx_of_first[tid] = i; err_of_first[tid] = err; }
                }
            }
        }
    };
    if (threads == 1) work(0);
    else {
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
        for (auto& th : pool) th.join();
    }
    int64_t best = INT64_MAX;
    int best_err = 0;
    for (int t = 0; t < threads; ++t)
        if (idx_of_first[t] < best) { best = idx_of_first[t]; best_err = err_of_first[t]; }
    if (best != INT64_MAX) { if (first_index) *first_index = best; if (first_errno) *first_errno = best_err; }
    return failed.load();
}

}  // extern "C"
