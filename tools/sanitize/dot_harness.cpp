#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include "../../include/nfact_b200.h"
namespace nfb { void set_error(const std::string&) {} }
int main(int argc, char** argv) {
    int ok = 0, bad = 0;
    for (int a = 1; a < argc; ++a) {
        FILE* f = fopen(argv[a], "rb"); if (!f) continue;
        fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
        char* src = (char*)malloc(n ? n : 1);
        if (fread(src, 1, n, f) != (size_t)n) { fclose(f); free(src); continue; }
        fclose(f);
        for (int nt : {1, 4}) {
            int64_t lines = 0;
            if (nfb_dot_count_lines(src, n, nt, &lines) != 0) { ++bad; continue; }
            double* r = (double*)malloc(sizeof(double) * (lines ? lines : 1));
            double* c = (double*)malloc(sizeof(double) * (lines ? lines : 1));
            double* v = (double*)malloc(sizeof(double) * (lines ? lines : 1));
            if (nfb_dot_parse(src, n, nt, lines, r, c, v) == 0) ++ok; else ++bad;
            if (lines >= 1) {
                int32_t* ri = (int32_t*)malloc(4 * (lines > 1 ? lines - 1 : 1));
                int32_t* ci = (int32_t*)malloc(4 * (lines > 1 ? lines - 1 : 1));
                double* vi = (double*)malloc(8 * (lines > 1 ? lines - 1 : 1));
                int64_t shape[2];
                if (nfb_dot_parse_coo(src, n, nt, lines, ri, ci, vi, shape) == 0) ++ok; else ++bad;
                free(ri); free(ci); free(vi);
            }
            free(r); free(c); free(v);
        }
        free(src);
    }
    printf("parsed %d rejected %d\n", ok, bad);
    return 0;
}
