// ASAN harness: decode every file given on the command line with exact-size heap buffers
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/nfact_b200.h"
namespace nfb { void set_error(const std::string&) {} }
int main(int argc, char** argv) {
    int ok = 0, bad = 0;
    for (int a = 1; a < argc; ++a) {
        FILE* f = fopen(argv[a], "rb"); if (!f) continue;
        fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
        unsigned char* src = (unsigned char*)malloc(n ? n : 1);          // exact size: ASAN sees any over-read
        if (fread(src, 1, n, f) != (size_t)n) { fclose(f); free(src); continue; }
        fclose(f);
        for (int nt : {1, 3}) {
            int64_t size = 0, wr = 0;
            if (nfb_lz4_frame_size(src, n, nt, &size) != 0) { ++bad; continue; }
            unsigned char* dst = (unsigned char*)malloc(size ? size : 1);  // exact size: any over-write is caught
            if (nfb_lz4_frame_decompress(src, n, dst, size, nt, &wr) == 0) ++ok; else ++bad;
            free(dst);
        }
        free(src);
    }
    printf("decoded %d rejected %d\n", ok, bad);
    return 0;
}
