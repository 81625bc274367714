import ctypes as C, zlib, numpy as np, sys
lib = C.CDLL("/tmp/asan/libinf_asan.so")
f = lib.llfe_inflate_zlib_mt
f.restype = C.c_int
f.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_int]
def run(z, cap, th):
    out = (C.c_ubyte * max(cap, 1))()
    got = C.c_size_t(0)
    rc = f(z, len(z), out, cap, C.byref(got), th)
    return rc, bytes(out[:got.value])
rng = np.random.default_rng(3)
walk = (np.cumsum(rng.integers(-2, 3, 2_500_000)) & 255).astype(np.uint8).tobytes()
noise = rng.integers(0, 256, 1_200_000, dtype=np.uint8).tobytes()
four = rng.integers(0, 4, 5_000_000, dtype=np.uint8).tobytes()
block = rng.integers(0, 256, 30000, dtype=np.uint8).tobytes()
rep = b"".join(block[:int(k)] + bytes([i & 255]) for i, k in enumerate(rng.integers(20000, 30000, 100)))
n = 0
for data in (walk, noise, four, rep, walk[:700000] + noise[:500000] + four[:900000]):
    for level, strat in ((1, 0), (6, 0), (6, zlib.Z_FILTERED), (6, zlib.Z_FIXED), (0, 0), (6, zlib.Z_HUFFMAN_ONLY)):
        c = zlib.compressobj(level, zlib.DEFLATED, 15, 8, strat)
        z = c.compress(data) + c.flush()
        for th in (2, 4, 8):
            for cap in (len(data), len(data) // 2 + 3, len(data) + 10):
                rc, out = run(z, cap, th)
                assert rc == 0 and out == data[:cap], (level, strat, th, cap)
                n += 1
        # damaged
        for _ in range(6):
            b = bytearray(z)
            b[int(rng.integers(2, len(b)))] ^= 1 << int(rng.integers(0, 8))
            for th in (3, 4):
                run(bytes(b), len(data), th)
        for cut in (len(z) // 3, len(z) - 3):
            run(z[:cut], len(data), 4)
print("asan run ok", n)
