// The product's RPO / RPX permutations (miden-vm_b200/csrc/rescue.cuh, __host__ __device__) compiled for the host:
//   test_rescue kat            Rpo256::hash_elements(&[0..=i]) for i = 0..18 (one line of 4 decimal felts each): the reference's
//                              known-answer table, built here from the permutation with AlgebraicSponge's padding rule
//   test_rescue perm <kind>    the permutation (3 = RPO, 4 = RPX) on 12 decimal felts read from stdin, repeated until EOF
#include "../../miden-vm_b200/csrc/rescue.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
int main(int argc, char** argv) {
    if (argc >= 2 && !strcmp(argv[1], "kat")) {
        for (int n = 1; n <= 19; n++) {
            unsigned long st[12] = {0};
            st[8] = n % 8;
            int i = 0;
            for (int k = 0; k < n; k++) { st[i++] = k; if (i == 8) { rsc::rpo_permute((gl::u64*)st); i = 0; } }
            if (i > 0) { while (i < 8) st[i++] = 0; rsc::rpo_permute((gl::u64*)st); }
            printf("%lu %lu %lu %lu\n", st[0], st[1], st[2], st[3]);
        }
        return 0;
    }
    if (argc >= 3 && !strcmp(argv[1], "perm")) {
        int kind = atoi(argv[2]);
        unsigned long st[12];
        for (;;) {
            for (int i = 0; i < 12; i++) if (scanf("%lu", &st[i]) != 1) return 0;
            if (kind == 4) rsc::rpx_permute((gl::u64*)st); else rsc::rpo_permute((gl::u64*)st);
            for (int i = 0; i < 12; i++) printf("%lu%c", st[i], i == 11 ? '\n' : ' ');
        }
    }
    return 2;
}
