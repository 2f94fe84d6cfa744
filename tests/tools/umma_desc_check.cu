// Host-side check that the hand-encoded UMMA descriptors and the shared-memory operand layout of
// fbpinns_b200/csrc/fbp_tc.cuh agree with the definitions in the CUTLASS/CuTe headers shipped in this image
// (cute/arch/mma_sm100_desc.hpp, cute/atom/mma_traits_sm100.hpp).  Prints "OK" and exits 0 on agreement.
// Built and run by tests/test_host_logic.py::test_umma_descriptors_match_cute (no GPU needed).
#include <cstdio>
#include <cute/tensor.hpp>
#include <cute/arch/mma_sm100_desc.hpp>
#include <cute/atom/mma_traits_sm100.hpp>
#include "fbp_tc.cuh"

using namespace cute;

int main() {
    int bad = 0;
    using T = tfloat32_t;
    // canonical K-major SWIZZLE_NONE layout of a 32 x 32 tf32 operand, K atoms innermost
    auto lay = tile_to_shape(UMMA::Layout_K_INTER_Atom<T>{}, Shape<_32, _32>{}, Step<_2, _1>{});
    for (int n = 0; n < 32; ++n)
        for (int k = 0; k < 32; ++k)
            if ((int)lay(n, k) != fbptc::bcore_index(n, k)) ++bad;
    if (bad) printf("layout: %d elements differ from the CuTe canonical layout\n", bad);
    // the strides make_umma_desc<Major::K> would put into the descriptor
    auto u128 = recast_layout<T, uint128_t>(lay.layout_b());
    auto canon = logical_divide(u128, Tile<Layout<_8, _1>, Layout<_2, _1>>{});
    const uint32_t sbo = (uint32_t)stride<0, 1>(canon) * 16, lbo = (uint32_t)stride<1, 0>(canon) * 16;
    if (sbo != fbptc::B_SBO || lbo != fbptc::B_LBO) { printf("strides: cute SBO %u LBO %u, ours %u %u\n", sbo, lbo, fbptc::B_SBO, fbptc::B_LBO); ++bad; }
    for (uint32_t addr : {0x0u, 0x12340u, 0x37ff0u}) {
        UMMA::SmemDescriptor d;
        d.version_ = 1; d.lbo_mode_ = 0; d.layout_type_ = uint8_t(UMMA::LayoutType::SWIZZLE_NONE); d.base_offset_ = 0;
        d.start_address_ = (uint16_t)(addr >> 4); d.leading_byte_offset_ = (uint16_t)(lbo >> 4); d.stride_byte_offset_ = (uint16_t)(sbo >> 4);
        const uint64_t ours = fbptc::make_smem_desc(addr, fbptc::B_LBO, fbptc::B_SBO);
        if (d.desc_ != ours) { printf("smem desc @%x: cute %016llx ours %016llx\n", addr, (unsigned long long)d.desc_, (unsigned long long)ours); ++bad; }
    }
    auto i16 = UMMA::make_instr_desc<T, T, float, 128, 16, UMMA::Major::K, UMMA::Major::K>();
    auto i32 = UMMA::make_instr_desc<T, T, float, 128, 32, UMMA::Major::K, UMMA::Major::K>();
    if (i16.desc_ != fbptc::make_idesc_tf32(128, 16) || i32.desc_ != fbptc::make_idesc_tf32(128, 32)) {
        printf("instr desc: cute %08x %08x ours %08x %08x\n", (unsigned)i16.desc_, (unsigned)i32.desc_, fbptc::make_idesc_tf32(128, 16), fbptc::make_idesc_tf32(128, 32));
        ++bad;
    }
    // operand images of the tensor-core weight gradient: the same canonical K-major form with a padded K step; with the
    // unpadded step it must coincide with the layout above, and the padded strides are what the descriptor carries
    for (int n = 0; n < 32; ++n)
        for (int k = 0; k < 32; ++k) {
            if (fbptc::gk_index(n, k, 128) != fbptc::bcore_index(n, k)) ++bad;
            if (fbptc::gk_index(n, k) != (n >> 3) * (int)(fbptc::GK_SBO / 4) + (k >> 2) * (int)(fbptc::GK_LBO / 4) + (n & 7) * 4 + (k & 3)) ++bad;
        }
    static_assert(fbptc::GK_LBO % 16 == 0 && fbptc::GK_SBO % 16 == 0, "descriptor strides are in units of 16 bytes");
    {   // 32 lanes storing one image row (32 consecutive points) must hit 32 different banks
        for (int row = 0; row < 32; ++row) {
            unsigned seen = 0;
            for (int k = 0; k < 32; ++k) seen |= 1u << (fbptc::gk_index(row, k) & 31);
            if (seen != 0xffffffffu) ++bad;
        }
    }
    {   // staging pass of the B operands (tc_stage_b): every W1[j][k] exactly once; the 32 stores of a warp hit 32 different banks
        // for B1[n = j][k] and at least 16 for B2[n = k][j]
        unsigned char seen[32][32] = {};
        for (int i = 0; i < 1024; ++i) seen[fbptc::stage_b_row(i)][fbptc::stage_b_col(i)]++;
        for (int j = 0; j < 32; ++j)
            for (int k = 0; k < 32; ++k)
                if (seen[j][k] != 1) ++bad;
        for (int w = 0; w < 32; ++w) {
            unsigned b1 = 0, b2 = 0;
            for (int l = 0; l < 32; ++l) {
                const int j = fbptc::stage_b_row(32 * w + l), k = fbptc::stage_b_col(32 * w + l);
                b1 |= 1u << (fbptc::bcore_index(j, k) & 31);
                b2 |= 1u << (fbptc::bcore_index(k, j) & 31);
            }
            if (b1 != 0xffffffffu || __builtin_popcount(b2) < 16) ++bad;
        }
    }
    // operand descriptors as base + constant (desc_advance) = the descriptor of the advanced address, over the whole shared memory
    for (uint32_t base : {0x400u, 0x5480u, 0x11480u})
        for (uint32_t off : {0u, 288u, 4608u, 73728u, 147440u}) {
            if (fbptc::desc_advance(fbptc::make_smem_desc(base, fbptc::GK_LBO, fbptc::GK_SBO), off) !=
                fbptc::make_smem_desc(base + off, fbptc::GK_LBO, fbptc::GK_SBO)) ++bad;
            if (fbptc::desc_advance(fbptc::make_smem_desc(base, fbptc::B_LBO, fbptc::B_SBO), off) !=
                fbptc::make_smem_desc(base + off, fbptc::B_LBO, fbptc::B_SBO)) ++bad;
        }
    printf(bad ? "MISMATCH\n" : "OK\n");
    return bad ? 1 : 0;
}
