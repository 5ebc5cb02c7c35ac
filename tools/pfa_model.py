"""Index-map model of the natural-order prime-factor slice FFT (k_slice.cu) + shared-memory conflict count.
Development aid: verifies the maps against numpy.fft and scores (SA, SB) pitches."""
import numpy as np
P1, P2, P3 = 43, 15, 14
N = P1 * P2 * P3
Q = P2 * P3
C1, C2, C3 = N // P1, N // P2, N // P3   # 210, 602, 645

def col_info(r):
    """column r of the 43-axis: (i2, i3, q) with base0 = (C2 i2 + C3 i3) % N = r + Q q"""
    inv2 = pow(C2 % P2, -1, P2); inv3 = pow(C3 % P3, -1, P3)
    i2 = (r * inv2) % P2; i3 = (r * inv3) % P3
    base0 = (C2 * i2 + C3 * i3) % N
    assert base0 % Q == r
    return i2, i3, base0 // Q

def fwd(x):
    L2 = np.zeros((P1, P2, P3), complex)
    s3 = C3 % P3
    assert s3 == 1
    for n0 in range(C3):
        n1, n2, s = n0 % P1, n0 % P2, n0 % P3
        v = np.array([x[n0 + C3 * ((c - s) % P3)] for c in range(P3)])
        L2[n1, n2, :] = np.fft.fft(v)
    L2 = np.fft.fft(L2, axis=1)
    Z = np.zeros(N, complex)
    for r in range(Q):
        k2, k3, q = col_info(r)
        X = np.fft.fft(L2[:, k2, k3])
        for k1 in range(P1):
            Z[r + Q * ((q + k1) % P1)] = X[k1]
    return Z

def inv(R):
    L2 = np.zeros((P1, P2, P3), complex)
    for r in range(Q):
        n2, n3, q = col_info(r)
        col = np.array([R[r + Q * ((q + n1) % P1)] for n1 in range(P1)])
        L2[:, n2, n3] = np.fft.ifft(col) * P1
    L2 = np.fft.ifft(L2, axis=1) * P2
    z = np.zeros(N, complex)
    for k0 in range(C3):
        k1, k2, s = k0 % P1, k0 % P2, k0 % P3
        v = np.fft.ifft(L2[k1, k2, :]) * P3
        for k3 in range(P3):
            z[k0 + C3 * ((k3 - s) % P3)] = v[k3]
    return z

def wavefronts(addrs):
    """addrs: list of 8-byte-unit addresses per lane (None = inactive) for one warp instruction"""
    tot = 0
    for h in range(0, 32, 16):
        banks = {}
        for a in addrs[h:h + 16]:
            if a is None: continue
            banks.setdefault(a % 16, set()).add(a)
        tot += max([len(v) for v in banks.values()], default=0)
    return tot

def score(SA, SB, NT=448):
    res = {}
    # pass A natural side (rotated), lanes <-> r, 224 slots per part
    wf = ideal = 0
    for n in range(P1):
        for w0 in range(0, 224, 32):
            ad = []
            for r in range(w0, w0 + 32):
                if r >= Q: ad.append(None); continue
                _, _, q = col_info(r); ad.append(r + Q * ((q + n) % P1))
            wf += wavefronts(ad); ideal += 2 if sum(a is not None for a in ad) > 16 else 1
    res['A_nat'] = wf / ideal
    wf = ideal = 0
    for k1 in range(P1):
        for w0 in range(0, 224, 32):
            ad = []
            for r in range(w0, w0 + 32):
                if r >= Q: ad.append(None); continue
                i2, i3, _ = col_info(r); ad.append(k1 * SA + i2 * SB + i3)
            wf += wavefronts(ad); ideal += 2 if sum(a is not None for a in ad) > 16 else 1
    res['A_l2'] = wf / ideal
    wf = ideal = 0
    for w0 in range(0, P1 * P3, 32):
        for n2 in range(P2):
            ad = []
            for t in range(w0, w0 + 32):
                if t >= P1 * P3: ad.append(None); continue
                ad.append((t % P1) * SA + n2 * SB + t // P1)
            wf += wavefronts(ad); ideal += 2 if sum(a is not None for a in ad) > 16 else 1
    res['B'] = wf / ideal
    wf = ideal = 0
    for w0 in range(0, C3, 32):
        for c in range(P3):
            ad = []
            for t in range(w0, w0 + 32):
                if t >= C3: ad.append(None); continue
                ad.append((t % P1) * SA + (t % P2) * SB + c)
            wf += wavefronts(ad); ideal += 2 if sum(a is not None for a in ad) > 16 else 1
    res['C'] = wf / ideal
    return res

if __name__ == "__main__":
    rs = np.random.RandomState(0)
    x = rs.randn(N) + 1j * rs.randn(N)
    print("fwd err", np.abs(fwd(x) - np.fft.fft(x)).max())
    print("inv err", np.abs(inv(x) - np.fft.ifft(x) * N).max())
    best = []
    for SB in range(14, 18):
        for SA in range(P2 * SB, P2 * SB + 8):
            s = score(SA, SB)
            best.append((max(s.values()), sum(s.values()), SA, SB, s))
    best.sort(key=lambda t: (t[1]))
    for b in best[:8]: print(b[2:], )
