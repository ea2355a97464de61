// io_helper.cpp -- files that may be compressed, by extension, like the reference's io_helper.rs
// (:18-48): ".gz" (multi-member gzip), ".bz2", ".zst"; anything else is plain bytes.
//
// The codecs are the system's shared libraries (libz.so.1, libbz2.so.1.0, libzstd.so.1), bound
// at run time with dlopen so that libfastrank_b200.so itself has no link-time dependency on them
// and still loads where one is missing -- opening such a file then fails with an error that names
// the library.  Host-side I/O only: nothing here is on the score -> rank -> metric path.
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

#include "host.hpp"

namespace frb {
namespace {

bool ends_with(const std::string &s, const char *ext) {
    const size_t n = strlen(ext);
    return s.size() >= n && s.compare(s.size() - n, n, ext) == 0;
}

void *open_lib(std::initializer_list<const char *> names, const char *what) {
    for (const char *nm : names) {
        if (void *h = dlopen(nm, RTLD_NOW | RTLD_LOCAL)) return h;
    }
    throw Error(std::string(what) + " support needs " + *names.begin() + ", which could not be loaded");
}

template <typename F>
F sym(void *h, const char *name) {
    void *p = dlsym(h, name);
    if (!p) throw Error(std::string("missing symbol ") + name);
    return (F)p;
}

[[noreturn]] void not_found(const std::string &path) {
    throw Error(path + ": Os { code: 2, kind: NotFound, message: \"No such file or directory\" }");
}

std::string read_plain(const std::string &path) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) not_found(path);
    std::string out;
    if (fseek(f, 0, SEEK_END) == 0) {
        const long size = ftell(f);
        if (size > 0) out.reserve((size_t)size);
        fseek(f, 0, SEEK_SET);
    }
    std::vector<char> buf((size_t)4 << 20);
    size_t n;
    while ((n = fread(buf.data(), 1, buf.size(), f)) > 0) out.append(buf.data(), n);
    fclose(f);
    return out;
}

// ---- gzip (zlib's gz* layer reads concatenated members like flate2's MultiGzDecoder) ----
struct Zlib {
    void *(*gzopen)(const char *, const char *);
    int (*gzread)(void *, void *, unsigned);
    int (*gzwrite)(void *, const void *, unsigned);
    int (*gzclose)(void *);
    static Zlib &get() {
        static Zlib z = [] {
            void *h = open_lib({"libz.so.1", "libz.so"}, "gzip");
            Zlib r;
            r.gzopen = sym<decltype(r.gzopen)>(h, "gzopen");
            r.gzread = sym<decltype(r.gzread)>(h, "gzread");
            r.gzwrite = sym<decltype(r.gzwrite)>(h, "gzwrite");
            r.gzclose = sym<decltype(r.gzclose)>(h, "gzclose");
            return r;
        }();
        return z;
    }
};

// ---- bzip2 (the library's stdio-like layer) ----
struct Bz2 {
    void *(*bzopen)(const char *, const char *);
    int (*bzread)(void *, void *, int);
    int (*bzwrite)(void *, void *, int);
    void (*bzclose)(void *);
    static Bz2 &get() {
        static Bz2 b = [] {
            void *h = open_lib({"libbz2.so.1.0", "libbz2.so.1", "libbz2.so"}, "bzip2");
            Bz2 r;
            r.bzopen = sym<decltype(r.bzopen)>(h, "BZ2_bzopen");
            r.bzread = sym<decltype(r.bzread)>(h, "BZ2_bzread");
            r.bzwrite = sym<decltype(r.bzwrite)>(h, "BZ2_bzwrite");
            r.bzclose = sym<decltype(r.bzclose)>(h, "BZ2_bzclose");
            return r;
        }();
        return b;
    }
};

// ---- zstd (streaming decoder: frames of unknown size, concatenated frames) ----
struct ZBuf {
    void *ptr;
    size_t size, pos;
};
struct Zstd {
    void *(*create)();
    size_t (*free_)(void *);
    size_t (*decompress)(void *, ZBuf *out, ZBuf *in);
    unsigned (*is_error)(size_t);
    size_t (*bound)(size_t);
    size_t (*compress)(void *, size_t, const void *, size_t, int);
    static Zstd &get() {
        static Zstd z = [] {
            void *h = open_lib({"libzstd.so.1", "libzstd.so"}, "zstd");
            Zstd r;
            r.create = sym<decltype(r.create)>(h, "ZSTD_createDStream");
            r.free_ = sym<decltype(r.free_)>(h, "ZSTD_freeDStream");
            r.decompress = sym<decltype(r.decompress)>(h, "ZSTD_decompressStream");
            r.is_error = sym<decltype(r.is_error)>(h, "ZSTD_isError");
            r.bound = sym<decltype(r.bound)>(h, "ZSTD_compressBound");
            r.compress = sym<decltype(r.compress)>(h, "ZSTD_compress");
            return r;
        }();
        return z;
    }
};

}  // namespace

std::string read_file_by_extension(const std::string &path) {  // io_helper.rs:18-29
    if (ends_with(path, ".gz")) {
        if (!std::ifstream(path)) not_found(path);
        Zlib &z = Zlib::get();
        void *f = z.gzopen(path.c_str(), "rb");
        if (!f) not_found(path);
        std::string out;
        std::vector<char> buf(1 << 20);
        int n;
        while ((n = z.gzread(f, buf.data(), (unsigned)buf.size())) > 0) out.append(buf.data(), (size_t)n);
        z.gzclose(f);
        if (n < 0) throw Error(path + ": corrupt gzip stream");
        return out;
    }
    if (ends_with(path, ".bz2")) {
        if (!std::ifstream(path)) not_found(path);
        Bz2 &b = Bz2::get();
        void *f = b.bzopen(path.c_str(), "rb");
        if (!f) not_found(path);
        std::string out;
        std::vector<char> buf(1 << 20);
        int n;
        while ((n = b.bzread(f, buf.data(), (int)buf.size())) > 0) out.append(buf.data(), (size_t)n);
        b.bzclose(f);
        if (n < 0) throw Error(path + ": corrupt bzip2 stream");
        return out;
    }
    if (ends_with(path, ".zst")) {
        const std::string raw = read_plain(path);
        Zstd &z = Zstd::get();
        void *st = z.create();
        if (!st) throw Error("zstd: cannot create a decoder");
        std::string out;
        std::vector<char> buf(1 << 20);
        ZBuf in{(void *)raw.data(), raw.size(), 0};
        size_t rc = 0;
        while (in.pos < in.size) {
            ZBuf o{buf.data(), buf.size(), 0};
            rc = z.decompress(st, &o, &in);
            if (z.is_error(rc)) {
                z.free_(st);
                throw Error(path + ": corrupt zstd stream");
            }
            out.append(buf.data(), o.pos);
        }
        for (;;) {  // what the decoder still holds back
            ZBuf o{buf.data(), buf.size(), 0};
            rc = z.decompress(st, &o, &in);
            if (z.is_error(rc)) break;
            out.append(buf.data(), o.pos);
            if (o.pos < o.size) break;
        }
        z.free_(st);
        return out;
    }
    return read_plain(path);
}

void write_file_by_extension(const std::string &path, const std::string &content) {  // io_helper.rs:31-48
    if (ends_with(path, ".gz")) {
        Zlib &z = Zlib::get();
        void *f = z.gzopen(path.c_str(), "wb");
        if (!f) throw Error("could not create " + path);
        size_t at = 0;
        while (at < content.size()) {
            const unsigned n = (unsigned)std::min<size_t>(content.size() - at, 1u << 30);
            if (z.gzwrite(f, content.data() + at, n) <= 0) {
                z.gzclose(f);
                throw Error("could not write " + path);
            }
            at += n;
        }
        if (z.gzclose(f) != 0) throw Error("could not write " + path);
        return;
    }
    if (ends_with(path, ".bz2")) {
        Bz2 &b = Bz2::get();
        void *f = b.bzopen(path.c_str(), "wb");
        if (!f) throw Error("could not create " + path);
        size_t at = 0;
        while (at < content.size()) {
            const int n = (int)std::min<size_t>(content.size() - at, 1u << 30);
            if (b.bzwrite(f, (void *)(content.data() + at), n) < 0) {
                b.bzclose(f);
                throw Error("could not write " + path);
            }
            at += (size_t)n;
        }
        b.bzclose(f);
        return;
    }
    std::string bytes;
    const std::string *payload = &content;
    if (ends_with(path, ".zst")) {
        Zstd &z = Zstd::get();
        bytes.resize(z.bound(content.size()));
        const size_t n = z.compress(&bytes[0], bytes.size(), content.data(), content.size(), 3);
        if (z.is_error(n)) throw Error("zstd: compression failed for " + path);
        bytes.resize(n);
        payload = &bytes;
    }
    std::ofstream out(path, std::ios::binary);
    if (!out) throw Error("could not create " + path);
    out.write(payload->data(), (std::streamsize)payload->size());
    if (!out) throw Error("could not write " + path);
}

}  // namespace frb
