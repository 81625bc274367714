#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <vector>
#include <thread>
extern "C" int llfe_inflate_zlib_mt(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_cap, size_t* out_len, int threads);
static std::vector<uint8_t> rd(const char* p) { FILE* f = fopen(p, "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET); std::vector<uint8_t> v(n); if (fread(v.data(), 1, n, f) != (size_t)n) abort(); fclose(f); return v; }
int main() {
    auto z = rd("/tmp/asan/z.bin"), ref = rd("/tmp/asan/ref.bin"), zb = rd("/tmp/asan/zbad.bin");
    int bad = 0;
    auto job = [&](int id) {
        std::vector<uint8_t> out(ref.size());
        for (int it = 0; it < 6; ++it) {
            size_t got = 0;
            int th = 2 + (it + id) % 5;
            int rc = llfe_inflate_zlib_mt(z.data(), z.size(), out.data(), out.size(), &got, th);
            if (rc != 0 || got != ref.size() || memcmp(out.data(), ref.data(), got)) __sync_fetch_and_add(&bad, 1);
            llfe_inflate_zlib_mt(zb.data(), zb.size(), out.data(), out.size(), &got, th);
            llfe_inflate_zlib_mt(z.data(), z.size(), out.data(), out.size() / 2, &got, th);
        }
    };
    std::thread a(job, 0), b(job, 1);
    a.join(); b.join();
    printf("tsan driver done, bad = %d\n", bad);
    return bad;
}
