"""NumPy model of the in-kernel real FFT (asr-featext-opencl_b200/csrc/afe_fft.cuh): same lane/slot index maps.

N2-point real FFT = M=N2/2 complex FFT of z[n]=x[2n]+i x[2n+1], M = 16*R (R = lanes per frame, 16 or 8):
  stage A  : lane n2 holds z[R*n1+n2], n1=0..15 -> radix-16 over n1 (4x4, output k1 at slot 4*(k1%4)+k1//4)
  twiddle  : * exp(-2 pi i n2 k1 / M)
  exchange : S[k1][n2] through shared memory
  stage B  : R-point FFT over n2 for each k1 -> Z[k1+16*k2]
  split    : X[k], X[M-k] from Z[k], Z[M-k]; k=0 -> X[0], X[M]; k=M/2 separately
Used by tests/test_host_logic.py to pin the decomposition against numpy.fft.rfft.
"""
import numpy as np


def fft4(a0, a1, a2, a3):
    s02, d02, s13, d13 = a0 + a2, a0 - a2, a1 + a3, a1 - a3
    return s02 + s13, d02 - 1j * d13, s02 - s13, d02 + 1j * d13


def fft16_slots(x):
    """x: list of 16 complex (natural order). Returns list y with y[4*(k%4)+k//4] = X[k]."""
    x = list(x)
    for b in range(4):
        x[b], x[4 + b], x[8 + b], x[12 + b] = fft4(x[b], x[4 + b], x[8 + b], x[12 + b])
    for c in range(1, 4):
        for b in range(1, 4):
            x[4 * c + b] = x[4 * c + b] * np.exp(-2j * np.pi * b * c / 16)
    for c in range(4):
        x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3] = fft4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3])
    return x


def fft8_slots(x):
    """y[2*(k%4)+k//4] = X[k]"""
    x = list(x)
    for b in range(2):
        x[b], x[2 + b], x[4 + b], x[6 + b] = fft4(x[b], x[2 + b], x[4 + b], x[6 + b])
    for c in range(1, 4):
        x[2 * c + 1] = x[2 * c + 1] * np.exp(-2j * np.pi * c / 8)
    for c in range(4):
        x[2 * c], x[2 * c + 1] = x[2 * c] + x[2 * c + 1], x[2 * c] - x[2 * c + 1]
    return x


def pos16(k):
    return 4 * (k % 4) + k // 4


def pos8(k):
    return 2 * (k % 4) + k // 4


def rfft_model(x):
    N2 = len(x)
    M = N2 // 2
    R = M // 16
    assert R in (8, 16)
    z = x[0::2] + 1j * x[1::2]
    # stage A + twiddle, per lane n2
    S = np.zeros((16, R), complex)
    for n2 in range(R):
        y = fft16_slots([z[R * n1 + n2] for n1 in range(16)])
        for k1 in range(16):
            S[k1, n2] = y[pos16(k1)] * np.exp(-2j * np.pi * n2 * k1 / M)
    # stage B
    Z = np.zeros(M, complex)
    for lane in range(R):
        for k1 in ([lane] if R == 16 else [lane, lane + 8]):
            if R == 16:
                y = fft16_slots(list(S[k1, :]))
                for k2 in range(16):
                    Z[k1 + 16 * k2] = y[pos16(k2)]
            else:
                y = fft8_slots(list(S[k1, :]))
                for k2 in range(8):
                    Z[k1 + 16 * k2] = y[pos8(k2)]
    # split; magnitudes of 2*X
    X = np.zeros(M + 1, complex)
    for lane in range(R):
        for m in range(8):
            k = lane + R * m
            a, b = Z[k], Z[(M - k) % M]
            w = np.exp(-2j * np.pi * k / N2)
            sr, si = a.real + b.real, a.imag - b.imag
            dr, di = a.real - b.real, a.imag + b.imag
            pr, pi = dr * w.real - di * w.imag, dr * w.imag + di * w.real
            X[k] = 0.5 * complex(sr + pi, si - pr)
            X[M - k] = 0.5 * complex(sr - pi, -(si + pr))
    X[M // 2] = np.conj(Z[M // 2])
    return X


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for N2 in (512, 256):
        x = rng.standard_normal(N2)
        err = np.abs(rfft_model(x) - np.fft.rfft(x)).max()
        print(N2, err)
        assert err < 1e-10
