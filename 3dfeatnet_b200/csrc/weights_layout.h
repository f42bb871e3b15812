// weights_layout.h -- the one definition of the packed (BN-folded) weight buffer shared by every
// forward kernel and by the host packer (exported through f3d_packed_weights_offsets).
//
// Order (each block starts on a 4-float = 16-byte boundary; W is (Cin,Cout) row-major, b is (Cout)):
//   detector   (models/feat3dnet.py:277-284,120-149): conv0 3->64, conv1 64->128, conv2 128->256,
//              conv_post_0 256->128, conv_post_1 128->64, attention 64->1, orientation 64->2
//   descriptor (models/feat3dnet.py:297-310,54-84):  conv0 3->32, conv1 32->64, conv_mid_0 128->MID
//              (rows 0..63 multiply the per-point features, rows 64..127 the tiled max-pool), conv_post_0 MID->F
#pragma once

namespace f3d {

enum WeightSlot {
    W_DET0 = 0, B_DET0, W_DET1, B_DET1, W_DET2, B_DET2, W_DETP0, B_DETP0, W_DETP1, B_DETP1, W_ATT, B_ATT, W_ORI, B_ORI,
    W_DESC0, B_DESC0, W_DESC1, B_DESC1, W_MID, B_MID, W_POST, B_POST, kNumWeightSlots
};

struct WeightLayout {
    int off[kNumWeightSlots];
    int size[kNumWeightSlots];
    int total;
    int mid;
    int feature_dim;
};

__host__ __device__ inline WeightLayout make_weight_layout(int feature_dim) {
    WeightLayout L;
    const int mid = feature_dim <= 64 ? 128 : 256;  // feat3dnet.py:300
    const int sz[kNumWeightSlots] = {3 * 64, 64, 64 * 128, 128, 128 * 256, 256, 256 * 128, 128, 128 * 64, 64, 64, 1, 128, 2,
                                     3 * 32, 32, 32 * 64, 64, 128 * mid, mid, mid * feature_dim, feature_dim};
    int o = 0;
    for (int i = 0; i < kNumWeightSlots; ++i) {
        L.off[i] = o;
        L.size[i] = sz[i];
        o += (sz[i] + 3) & ~3;
    }
    L.total = o;
    L.mid = mid;
    L.feature_dim = feature_dim;
    return L;
}

}  // namespace f3d
