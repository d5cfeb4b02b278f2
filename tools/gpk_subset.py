"""fixture tooling: keep every STEP-th genome of a .gpk (tools/mkdump.cpp format) so that tools/readgen can draw reads of a
large collection from a small file (every kept genome is still a genome of the indexed collection).
    python tools/gpk_subset.py IN.gpk OUT.gpk STEP"""
import struct
import sys

import numpy as np

src, dst, step = sys.argv[1], sys.argv[2], int(sys.argv[3])
raw = np.fromfile(src, dtype=np.uint8)
assert raw[:8].tobytes() == b"FGPK1\0\0\0"
nc, total = struct.unpack_from("<QQ", raw, 8)
p = 24
contigs = []
for _ in range(nc):
    g, ln = struct.unpack_from("<IQ", raw, p)
    contigs.append((g, ln))
    p += 12
packed = raw[p:p + (total + 3) // 4]
p += (total + 3) // 4
ne, = struct.unpack_from("<Q", raw, p)
exc = np.frombuffer(raw, dtype="<u8", count=ne, offset=p + 8)
codes = np.zeros(total, dtype=np.uint8)
for s in range(4):
    codes[s::4] = (packed[: (total - s + 3) // 4] >> (2 * s)) & 3
keep, kept_contigs, new_exc, pos, out_pos = [], [], [], 0, 0
for g, ln in contigs:
    if g % step == 0:
        keep.append(codes[pos:pos + ln])
        kept_contigs.append((g, ln))
        e = exc[(exc >= pos) & (exc < pos + ln)]
        new_exc.extend((e - pos + out_pos).tolist())
        out_pos += ln
    pos += ln
codes = np.concatenate(keep) if keep else np.zeros(0, dtype=np.uint8)
pad = (-codes.size) % 4
c4 = np.concatenate([codes, np.zeros(pad, dtype=np.uint8)]).reshape(-1, 4)
out_packed = (c4[:, 0] | (c4[:, 1] << 2) | (c4[:, 2] << 4) | (c4[:, 3] << 6)).astype(np.uint8)
with open(dst, "wb") as f:
    f.write(b"FGPK1\0\0\0")
    f.write(struct.pack("<QQ", len(kept_contigs), codes.size))
    for g, ln in kept_contigs:
        f.write(struct.pack("<IQ", g, ln))
    f.write(out_packed.tobytes())
    f.write(struct.pack("<Q", len(new_exc)))
    f.write(np.asarray(new_exc, dtype="<u8").tobytes())
print(f"{dst}: {len(kept_contigs)} contigs, {codes.size} bases")
