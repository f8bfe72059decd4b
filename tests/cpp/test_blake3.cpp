// The product's BLAKE3 (miden-vm_b200/csrc/blake3.cuh, __host__ __device__) compiled for the host:
//   test_blake3 <n_words>   hashes the 4*n bytes i mod 251 and prints the digest (tests/test_blake3.py compares it with
//   vectors made by the official crate's bindings).
#include "../../miden-vm_b200/csrc/blake3.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
int main(int argc, char** argv) {
    if (argc < 2) return 2;
    size_t nw = strtoull(argv[1], nullptr, 10);
    std::vector<unsigned char> bytes(4 * nw);
    for (size_t i = 0; i < bytes.size(); i++) bytes[i] = (unsigned char)(i % 251);
    b3::Hasher h; h.init();
    for (size_t i = 0; i < nw; i++) { unsigned w; memcpy(&w, &bytes[4 * i], 4); h.push(w); }
    unsigned out[8]; h.finish(out);
    for (int i = 0; i < 8; i++) for (int k = 0; k < 4; k++) printf("%02x", (out[i] >> (8 * k)) & 0xff);
    printf("\n");
    return 0;
}
