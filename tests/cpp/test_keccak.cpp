// The product's Keccak (miden-vm_b200/csrc/keccak.cuh, __host__ __device__) compiled for the host:
//   test_keccak hash <pad> <n_words>   Keccak/SHA3-256 of the 8*n bytes i mod 251 (pad = 1: Keccak-256, 6: SHA3-256)
//   test_keccak perm                   the permutation applied to the state 0, 1, ..., 24 (25 hex words)
#include "../../miden-vm_b200/csrc/keccak.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
int main(int argc, char** argv) {
    if (argc >= 2 && !strcmp(argv[1], "perm")) {
        unsigned long st[25];
        for (int i = 0; i < 25; i++) st[i] = (unsigned long)i;
        kk::permute((gl::u64*)st);
        for (int i = 0; i < 25; i++) printf("%016lx ", st[i]);
        printf("\n");
        return 0;
    }
    if (argc < 4) return 2;
    unsigned long pad = strtoul(argv[2], nullptr, 10);
    size_t nw = strtoull(argv[3], nullptr, 10);
    std::vector<unsigned char> bytes(8 * nw);
    for (size_t i = 0; i < bytes.size(); i++) bytes[i] = (unsigned char)(i % 251);
    kk::Hash256 h; h.init();
    for (size_t i = 0; i < nw; i++) { gl::u64 w; memcpy(&w, &bytes[8 * i], 8); h.push64(w); }
    gl::u64 out[4]; h.finish(out, pad);
    for (int i = 0; i < 4; i++) for (int k = 0; k < 8; k++) printf("%02x", (unsigned)((out[i] >> (8 * k)) & 0xff));
    printf("\n");
    return 0;
}
