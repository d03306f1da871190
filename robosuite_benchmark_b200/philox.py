"""Philox4x32-10 on the host (plain Python ints) -- the counter-based generator every random draw of the device path uses
(reset noise, object placement, synthetic actions, replay indices, policy noise).  Key/counter conventions: csrc/rsb_dev.h
(`rsb_philox`) and csrc/rsb_sac.cu."""
M0, M1, W0, W1, MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF


def philox4x32(counter, key):
    c = [int(x) & MASK for x in counter]
    k0, k1 = int(key[0]) & MASK, int(key[1]) & MASK
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c[3] ^ k1) & MASK, p0 & MASK]
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c
