#include "agc_reader.hpp"

#include <dlfcn.h>
#include <libgen.h>
#include <unistd.h>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

namespace pgrb200 {

namespace {

// the C API of libagc (agc/src/lib-cxx/agc-api.h:95-182), bound at run time
struct AgcApi {
    void *dl = nullptr;
    void *(*open)(char *, int) = nullptr;
    int (*close)(void *) = nullptr;
    int (*get_ctg_len)(const void *, const char *, const char *) = nullptr;
    int (*get_ctg_seq)(const void *, const char *, const char *, int, int, char *) = nullptr;
    int (*n_sample)(const void *) = nullptr;
    int (*n_ctg)(const void *, const char *) = nullptr;
    char **(*list_sample)(const void *, int *) = nullptr;
    char **(*list_ctg)(const void *, const char *, int *) = nullptr;
    int (*list_destroy)(char **) = nullptr;
    std::string err;
};

AgcApi &api() {
    static AgcApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        std::vector<std::string> cand;
        if (const char *e = getenv("PGR_B200_LIBAGC")) cand.push_back(e);
        char exe[4096];
        const ssize_t n = readlink("/proc/self/exe", exe, sizeof exe - 1);
        if (n > 0) { exe[n] = 0; cand.push_back(std::string(dirname(exe)) + "/libagc_ref.so"); }
        cand.push_back("libagc_ref.so");
        for (auto &c : cand) { a.dl = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL); if (a.dl) break; }
        if (!a.dl) { a.err = "libagc_ref.so not found (build it with `make -C pgr_tk_b200/host agc` where the reference's agc/ sources are, or set PGR_B200_LIBAGC)"; return; }
        auto sym = [&](const char *nm) { void *p = dlsym(a.dl, nm); if (!p) a.err = std::string("libagc lacks ") + nm; return p; };
        a.open = (void *(*)(char *, int))sym("agc_open");
        a.close = (int (*)(void *))sym("agc_close");
        a.get_ctg_len = (int (*)(const void *, const char *, const char *))sym("agc_get_ctg_len");
        a.get_ctg_seq = (int (*)(const void *, const char *, const char *, int, int, char *))sym("agc_get_ctg_seq");
        a.n_sample = (int (*)(const void *))sym("agc_n_sample");
        a.n_ctg = (int (*)(const void *, const char *))sym("agc_n_ctg");
        a.list_sample = (char **(*)(const void *, int *))sym("agc_list_sample");
        a.list_ctg = (char **(*)(const void *, const char *, int *))sym("agc_list_ctg");
        a.list_destroy = (int (*)(char **))sym("agc_list_destroy");
    });
    return a;
}

}  // namespace

AgcFile::~AgcFile() {}

bool AgcFile::open(const std::string &path, bool prefetching, std::string &err) {
    AgcApi &a = api();
    if (!a.err.empty()) { err = a.err; return false; }
    if (access(path.c_str(), R_OK) != 0) { err = "cannot open " + path; return false; }
    path_ = path; prefetching_ = prefetching;
    std::vector<char> fn(path.begin(), path.end());
    fn.push_back(0);
    void *h = a.open(fn.data(), 1);                                       // agc_io.rs:77-80
    if (!h) { err = "agc_open failed on " + path; return false; }
    int ns = a.n_sample(h);
    char **samples = a.list_sample(h, &ns);
    for (int i = 0; i < ns; i++) {
        int nc = a.n_ctg(h, samples[i]);
        char **ctgs = a.list_ctg(h, samples[i], &nc);
        for (int j = 0; j < nc; j++) {
            AgcContig c;
            c.sample = samples[i]; c.name = ctgs[j];
            c.len = (size_t)a.get_ctg_len(h, samples[i], ctgs[j]);
            ctgs_.push_back(std::move(c));
        }
        a.list_destroy(ctgs);
    }
    a.list_destroy(samples);
    a.close(h);
    return true;
}

bool AgcFile::fetch(size_t i0, size_t i1, int n_threads, std::vector<std::vector<uint8_t>> &out, std::string &err) {
    AgcApi &a = api();
    if (!a.err.empty()) { err = a.err; return false; }
    out.assign(i1 - i0, {});
    std::atomic<size_t> next{i0};
    std::atomic<bool> failed{false};
    auto work = [&]() {
        std::vector<char> fn(path_.begin(), path_.end());
        fn.push_back(0);
        void *h = a.open(fn.data(), prefetching_ ? 1 : 0);                // one handle per thread (agc_io.rs:262-271)
        if (!h) { failed.store(true); return; }
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= i1) break;
            const AgcContig &c = ctgs_[i];
            std::vector<uint8_t> buf(c.len + 2);
            // the reference asks for [0, len] into a buffer of len + 1 bytes and keeps len bytes (agc_io.rs:283-300)
            a.get_ctg_seq(h, c.sample.c_str(), c.name.c_str(), 0, (int)c.len, (char *)buf.data());
            buf.resize(c.len);
            out[i - i0] = std::move(buf);
        }
        a.close(h);
    };
    n_threads = std::max(1, std::min<int>(n_threads, (int)std::max<size_t>(1, i1 - i0)));
    std::vector<std::thread> th;
    for (int t = 1; t < n_threads; t++) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
    if (failed.load()) { err = "agc_open failed in a reader thread"; return false; }
    return true;
}

}  // namespace pgrb200
