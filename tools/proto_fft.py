"""numpy prototype of the exact lane/register data flow of the CUDA fbank kernel
(16 lanes x 16 complex registers, 256-point complex FFT of the packed real frame,
shuffle pairing with the lane-0 fix-up, real split, power).  Used to pin the index math
before writing CUDA; not part of the product."""
import numpy as np

def dft16_regs(v):
    """v: list of 16 complex (natural order) -> natural order, via 4x4 with static indices."""
    w16 = np.exp(-2j*np.pi*np.arange(16)/16)
    B = [[None]*4 for _ in range(4)]  # B[na][klo4]
    for na in range(4):
        x0, x1, x2, x3 = v[na], v[na+4], v[na+8], v[na+12]
        s0, d0, s1, d1 = x0+x2, x0-x2, x1+x3, x1-x3
        B[na][0] = s0+s1; B[na][2] = s0-s1; B[na][1] = d0 - 1j*d1; B[na][3] = d0 + 1j*d1
    out = [None]*16
    for kl in range(4):
        y = [B[na][kl]*w16[(na*kl) % 16] for na in range(4)]
        s0, d0, s1, d1 = y[0]+y[2], y[0]-y[2], y[1]+y[3], y[1]-y[3]
        out[kl+0] = s0+s1; out[kl+8] = s0-s1; out[kl+4] = d0 - 1j*d1; out[kl+12] = d0 + 1j*d1
    return out

def frame_power(y512):
    """y512: windowed, zero padded real frame (512,). Returns P[0..255] = |X[k]|^2 via the lane flow."""
    z = y512[0::2] + 1j*y512[1::2]           # 256 complex
    W256 = np.exp(-2j*np.pi*np.arange(256)/256)
    # pass 1: lane n1 holds z[n1+16*n2]
    Y = np.zeros((16, 16), complex)          # [n1][klo]
    for n1 in range(16):
        v = [z[n1+16*n2] for n2 in range(16)]
        o = dft16_regs(v)
        for kl in range(16):
            Y[n1, kl] = o[kl]*W256[(n1*kl) % 256]
    # exchange; pass 2: lane klo reads Y[:, klo]
    Z = np.zeros((16, 16), complex)          # [lane=klo][r=khi]
    for kl in range(16):
        o = dft16_regs([Y[n1, kl] for n1 in range(16)])
        for r in range(16):
            Z[kl, r] = o[r]
    # check Z against fft
    ref = np.fft.fft(z)
    for kl in range(16):
        for r in range(16):
            assert abs(Z[kl, r]-ref[kl+16*r]) < 1e-9*abs(ref).max()
    # pairing: every lane handles own r=0..7 with partner lane (16-l)&15 register 15-r
    P = np.zeros(256)
    W512 = np.exp(-2j*np.pi*np.arange(512)/512)
    for l in range(16):
        pl = (16-l) & 15
        recv = [Z[pl, 15-r] for r in range(8)]     # generic shuffle result
        if l == 0:                                 # lane-0 fix-up: partner of 16r is 16(16-r)
            fixed = [Z[0, 0]] + [recv[r-1] for r in range(1, 8)]
            recv = fixed
        for r in range(8):
            k = l + 16*r
            A = Z[l, r]; Bc = np.conj(recv[r])
            S = A + Bc; D = A - Bc
            T = (-1j*W512[k])*D
            Xk = 0.5*(S+T)            # X[k]
            Xm = 0.5*np.conj(S-T)     # X[256-k]
            P[k] = abs(Xk)**2
            if k != 0:
                P[256-k] = abs(Xm)**2
        if l == 0:
            P[128] = abs(Z[0, 8])**2
    return P

rng = np.random.default_rng(0)
y = np.zeros(512); y[:400] = rng.standard_normal(400)
P = frame_power(y)
ref = np.abs(np.fft.rfft(y))**2
print("max rel err", np.max(np.abs(P-ref[:256])/ref[:256].max()))
assert np.allclose(P, ref[:256], rtol=1e-9, atol=1e-9*ref.max())
print("OK")
