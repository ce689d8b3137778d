"""Dev-time check (needs /root/reference): the VVC 64-point DCT-II integer matrix used by the
reference (transformer.rs:934-1234) is fully determined by 65 magnitudes c[j] ~ 64*sqrt(2)*cos(j*pi/128):
M[k][n] = sign * c[fold(k*(2n+1) mod 256)], M[0][n] = 64.  Prints c[] so it can be pasted as a standard constant."""
import re, sys
src = open('/root/reference/src/transformer.rs').read()
i = src.index('const TRANS_MATRIX_0_: [[i16; 32]; 64] = [')
j = src.index('\n];', i)
nums = [int(v) for v in re.findall(r'-?\d+', src[i + len('const TRANS_MATRIX_0_: [[i16; 32]; 64] = ['):j])]
assert len(nums) == 64 * 32, len(nums)
rows = [nums[r * 32:(r + 1) * 32] for r in range(64)]
M = []
for n, r in enumerate(rows):
    sign = 1 - 2 * (n & 1)
    M.append(r + [r[31 - q] * sign for q in range(32)])
c = [None] * 65
def fold(m):
    m %= 256
    s = 1
    if m > 128:
        m = 256 - m
    if m > 64:
        m = 128 - m
        s = -1
    return m, s
ok = True
for k in range(1, 64):
    for n in range(64):
        m, s = fold(k * (2 * n + 1))
        v = M[k][n] * s
        if c[m] is None:
            c[m] = v
        elif c[m] != v:
            ok = False
            print('mismatch', k, n, m, c[m], v)
c[0] = 91  # never used for k>=1 (k*(2n+1) mod 256 != 0); row 0 is the constant 64
print('structure ok:', ok)
print(c)
