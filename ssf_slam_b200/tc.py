"""Host-side preparation for the tcgen05 (tensor-core) kernels: fp32 -> (hi, lo) TF32 split and the shared-memory
operand image the UMMA descriptors expect (see csrc/tc_common.cuh)."""
import torch


def split_tf32(w):
    """hi = round-to-nearest(ties away) to 10 mantissa bits (== cvt.rna.tf32.f32), lo = w - hi (exact in fp32)."""
    bits = w.contiguous().view(torch.int32)
    hi = ((bits + 0x1000) & ~0x1FFF).view(torch.float32)
    return hi, w - hi


def kmajor_image(w):
    """w [R, K] (row r, K contiguous) -> flat fp32 tensor in the no-swizzle K-major core-matrix order:
    byte(r, k) = (k/4)*(R/8)*128 + (r/8)*128 + (r%8)*16 + (k%4)*4."""
    R, K = w.shape
    assert R % 8 == 0 and K % 4 == 0
    return w.view(R // 8, 8, K // 4, 4).permute(2, 0, 1, 3).contiguous().view(-1)


def weight_image(w):
    """w [N, K] (the reference's [Cout, Cin] conv weight) -> (hi image, lo image)."""
    hi, lo = split_tf32(w.float())
    return kmajor_image(hi), kmajor_image(lo)


def cost_volume_tc_pack(w2a, w2w, w3a, w3b, wn1, wn2, b2a, b2w, b3b, bn1, bn2, wn3, w3d, bn3):
    """Weight blob + parameter block of ``ssf_cost_volume_tc`` (csrc/cost_volume_tc.cu) for m = 64.
    Matrices are in the reference's [Cout, Cin] orientation (BatchNorm already folded); ``w3d`` is [3, m] K-major.
    Blob = six chunks [hi image | lo image]: mlp_convs[1], mlp_convs2[1], mlp_convs3[0][:, :m], mlp_convs3[1],
    weightnet1[0], weightnet1[3]; params = b2a b2w b3b bn1 bn2 wn3 W3d bn3 (fp32)."""
    chunks = []
    for w in (w2a, w2w, w3a, w3b, wn1, wn2):
        hi, lo = weight_image(w.contiguous())
        chunks += [hi, lo]
    blob = torch.cat(chunks).contiguous()
    params = torch.cat([b2a, b2w, b3b, bn1, bn2, wn3.reshape(-1), w3d.reshape(-1), torch.tensor([bn3, 0.0, 0.0, 0.0])]).float()
    assert blob.numel() * 4 == 5 * 32768 + 16384 and params.numel() == 516
    return blob, params.contiguous()


def dense_image(w):
    """w [N, K] (reference [Cout, Cin] orientation, BatchNorm folded) -> weight image of ``ssf_dense_tc``:
    for every 256-column tile, for every 32-wide K chunk: [hi image | lo image] of the [Nt, 32] block."""
    w = w.float().contiguous()
    N, K = w.shape
    assert K % 32 == 0 and N % 32 == 0 and (N <= 256 or N % 256 == 0)
    hi, lo = split_tf32(w)
    parts = []
    for n0 in range(0, N, 256):
        n1 = min(N, n0 + 256)
        for k0 in range(0, K, 32):
            parts += [kmajor_image(hi[n0:n1, k0:k0 + 32].contiguous()), kmajor_image(lo[n0:n1, k0:k0 + 32].contiguous())]
    return torch.cat(parts).contiguous()
