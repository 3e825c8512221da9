"""Developer probe driver (GPU box): which shared-memory layouts / descriptor encodings does
tcgen05.mma.kind::tf32 accept for K-major and MN-major operands?  Prints max |D - A B^T| per variant."""
import ctypes
import os
import numpy as np
import torch

lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libumma_probe.so"))
lib.umma_probe.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_uint64,
                           ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                           ctypes.c_void_p]
M, N, K = 128, 32, 8


def swz(addr, B, Mb, S):
    return addr ^ (((addr >> (Mb + S)) & ((1 << B) - 1)) << Mb)


def desc_hi(lbo, sbo, layout):
    return ((lbo >> 4) << 16) | ((sbo >> 4) << 32) | (1 << 46) | (layout << 61)


def idesc(a_mn, b_mn):
    return (1 << 4) | (2 << 7) | (2 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def image(mat, addr_fn, size):
    img = np.zeros(size // 4, np.float32)
    for r in range(mat.shape[0]):
        for k in range(mat.shape[1]):
            a = addr_fn(r, k)
            assert a % 4 == 0 and a < size, (r, k, a)
            img[a // 4] = mat[r, k]
    return img


def kmajor_noswz(rows):                 # [k-core j][row][16B], SBO=128, LBO=rows*16
    lbo = rows * 16
    return (lambda r, k: (k // 4) * lbo + r * 16 + (k % 4) * 4), lbo, 128, 0, rows * 16 * 2


def mn_noswz(rows):                     # [piece mi][8 k][16B]: SBO=128 between pieces, one k-block
    return (lambda r, k: (r // 4) * 128 + (k % 8) * 16 + (r % 4) * 4), rows * 32, 128, 0, rows * 32


def mn_noswz_swapped(rows):
    f, lbo, sbo, lay, size = mn_noswz(rows)
    return f, sbo, lbo, lay, size


def mn_sw128(rows):                     # 128B rows of 32 MN elements, 8 k rows per 1024B atom; atoms along MN at LBO
    lbo = 1024
    return (lambda r, k: swz((r // 32) * lbo + (k % 8) * 128 + (r % 32) * 4, 3, 4, 3)), lbo, 1024, 2, (rows // 32) * 1024


def mn_sw128_base32(rows):              # Swizzle<2,5,2>, k atom of 4?
    lbo = 1024
    return (lambda r, k: swz((r // 32) * lbo + (k % 8) * 128 + (r % 32) * 4, 2, 5, 2)), lbo, 1024, 1, (rows // 32) * 1024


def mn_base32(rows):                    # Layout_MN_SW128_32B_Atom: 128B rows (32 MN elems) x 4 k rows, Swizzle<2,5,2>
    lbo = 512                           # between MN atoms; k atoms (4 k each) at SBO
    sbo = (rows // 32) * 512
    return (lambda r, k: swz((k // 4) * sbo + (r // 32) * lbo + (k % 4) * 128 + (r % 32) * 4, 2, 5, 2)), lbo, sbo, 1, 2 * sbo


def mn_base32_b(rows):                  # same atom, k atoms contiguous (SBO=512), MN atoms at LBO=1024
    lbo, sbo = 1024, 512
    return (lambda r, k: swz((k // 4) * sbo + (r // 32) * lbo + (k % 4) * 128 + (r % 32) * 4, 2, 5, 2)), lbo, sbo, 1, (rows // 32) * 1024


def run(name, a_fn, b_fn, a_mn, b_mn):
    rng = np.random.RandomState(0)
    A = rng.randint(-4, 5, (M, K)).astype(np.float32)
    B = rng.randint(-4, 5, (N, K)).astype(np.float32)
    fa, lbo_a, sbo_a, lay_a, size_a = a_fn(M)
    fb, lbo_b, sbo_b, lay_b, size_b = b_fn(N)
    ia = torch.from_numpy(image(A, fa, size_a)).cuda()
    ib = torch.from_numpy(image(B, fb, size_b)).cuda()
    out = torch.zeros(M, N, device="cuda")
    rc = lib.umma_probe(ia.data_ptr(), ib.data_ptr(), size_a, size_b, desc_hi(lbo_a, sbo_a, lay_a),
                        desc_hi(lbo_b, sbo_b, lay_b), idesc(a_mn, b_mn), 1, 0, 0, N, out.data_ptr())
    want = A @ B.T
    err = np.abs(out.cpu().numpy() - want).max()
    print("%-46s rc=%d max|err|=%g  (|D|max %g, out max %g)" % (name, rc, err, np.abs(want).max(),
                                                               float(out.abs().max())))


def run_k32_sw128():
    """K-major SWIZZLE_128B for both operands, 4 MMAs of K=8 walking one 128-byte row (start += 32 B)"""
    K32 = 32
    rng = np.random.RandomState(1)
    A = rng.randint(-4, 5, (M, K32)).astype(np.float32)
    B = rng.randint(-4, 5, (N, K32)).astype(np.float32)
    fa = lambda r, k: swz((r // 8) * 1024 + (r % 8) * 128 + k * 4, 3, 4, 3)
    ia = torch.from_numpy(image(A, fa, M * 128)).cuda()
    ib = torch.from_numpy(image(B, fa, N * 128)).cuda()
    for lbo in (16, 1024):
        out = torch.zeros(M, N, device="cuda")
        rc = lib.umma_probe(ia.data_ptr(), ib.data_ptr(), M * 128, N * 128, desc_hi(lbo, 1024, 2), desc_hi(lbo, 1024, 2),
                            idesc(0, 0), 4, 32, 32, N, out.data_ptr())
        want = A @ B.T
        print("K-major SW128 both, 4 MMAs step 32B, lbo=%d: rc=%d max|err|=%g (|D|max %g)" %
              (lbo, rc, np.abs(out.cpu().numpy() - want).max(), np.abs(want).max()))


if __name__ == "__main__":
    run_k32_sw128()
    run("A MN base32 [katom][mi], B K-major", mn_base32, kmajor_noswz, 1, 0)
    run("A MN base32 [mi][katom], B K-major", mn_base32_b, kmajor_noswz, 1, 0)
    run("A K-major, B MN base32 [katom][ni]", kmajor_noswz, mn_base32, 0, 1)
    run("A MN base32, B MN base32", mn_base32, mn_base32, 1, 1)
    run("A MN base32_b, B MN base32_b", mn_base32_b, mn_base32_b, 1, 1)
    run("A K-major noswz, B K-major noswz (sanity)", kmajor_noswz, kmajor_noswz, 0, 0)
    run("A MN noswz (sbo=128), B K-major", mn_noswz, kmajor_noswz, 1, 0)
    run("A MN noswz swapped fields, B K-major", mn_noswz_swapped, kmajor_noswz, 1, 0)
    run("A MN sw128, B K-major", mn_sw128, kmajor_noswz, 1, 0)
    run("A MN sw128_base32, B K-major", mn_sw128_base32, kmajor_noswz, 1, 0)
    run("A K-major, B MN noswz", kmajor_noswz, mn_noswz, 0, 1)
    run("A K-major, B MN noswz swapped", kmajor_noswz, mn_noswz_swapped, 0, 1)
    run("A K-major, B MN sw128", kmajor_noswz, mn_sw128, 0, 1)
    run("A MN noswz, B MN noswz", mn_noswz, mn_noswz, 1, 1)
    run("A MN sw128, B MN sw128", mn_sw128, mn_sw128, 1, 1)
