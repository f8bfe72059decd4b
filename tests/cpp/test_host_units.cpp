// Command-line access to the product's host-side index logic (miden-vm_b200/csrc/host_transcript.hpp) so that
// tests/test_reference_vectors.py can run the reference's own TreeIndices unit vectors
// (crates/lifted-stark/src/lmcs/tree_indices.rs:245-370) against it without a GPU.
//   test_host_units fold <depth> <target> idx...      -> the folded, sorted, unique indices
//   test_host_units siblings <depth> idx...           -> "depth:position" of every missing sibling, emission order
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../miden-vm_b200/csrc/host_transcript.hpp"

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    unsigned depth = (unsigned)atoi(argv[2]);
    if (!strcmp(argv[1], "fold")) {
        if (argc < 4) return 2;
        unsigned target = (unsigned)atoi(argv[3]);
        std::vector<size_t> v;
        for (int i = 4; i < argc; i++) v.push_back((size_t)strtoull(argv[i], nullptr, 10));
        hostfs::Indices t = hostfs::Indices::make(v, depth).folded(target);
        for (size_t x : t.idx) printf("%zu ", x);
        printf("\n");
        return 0;
    }
    if (!strcmp(argv[1], "siblings")) {
        std::vector<size_t> v;
        for (int i = 3; i < argc; i++) v.push_back((size_t)strtoull(argv[i], nullptr, 10));
        for (auto& ds : hostfs::missing_siblings(hostfs::Indices::make(v, depth))) printf("%u:%zu ", ds.first, ds.second);
        printf("\n");
        return 0;
    }
    return 2;
}
