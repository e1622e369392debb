#!/usr/bin/env python
"""Derive the fp32 polynomial coefficients used by the packed sincos in csrc/ikl_packed.cuh and measure the
scheme's accuracy with an emulated fp32 pipeline (every op rounded to float32; FMA = one rounding).

  python tools/fit_sincos.py

Scheme (mod pi): k = rint(x/pi) (magic-number rounding), r = x - k*pi in three FMAs (Cody-Waite split of pi),
sin x = (-1)^k * (r + r*z*Ps(z)),  cos x = (-1)^k * (1 + z*Pc(z)),  z = r*r,  r in [-pi/2, pi/2].
Coefficients: weighted least squares on Chebyshev nodes in float64 (relative weight for sin), then rounded to fp32.
"""
import numpy as np

f32 = np.float32


def fma(a, b, c):
  return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def split_pi():
  import mpmath as mp
  mp.mp.prec = 200
  pi = mp.pi
  hi = f32(float(pi)); hi = f32(np.frombuffer(np.uint32(np.frombuffer(hi.tobytes(), np.uint32)[0] & 0xFFFFF000).tobytes(), f32)[0])
  mid = f32(float(pi - mp.mpf(float(hi)))); mid = f32(np.frombuffer(np.uint32(np.frombuffer(mid.tobytes(), np.uint32)[0] & 0xFFFFF000).tobytes(), f32)[0])
  lo = f32(float(pi - mp.mpf(float(hi)) - mp.mpf(float(mid))))
  return hi, mid, lo


def fit(deg_s, deg_c, half_width):
  n = 4000
  t = np.cos(np.pi * (np.arange(n) + 0.5) / n)          # Chebyshev nodes in [-1,1]
  r = (t + 1) / 2 * half_width
  r = r[r > 1e-6]
  z = r * r
  # sin(r) = r + r*z*Ps(z)  ->  Ps(z) = (sin(r)/r - 1)/z ; weight so that the ABSOLUTE error of sin is minimised
  ys = (np.sin(r) / r - 1) / z
  A = np.vander(z, deg_s, increasing=True)
  ws = r * z
  cs = np.linalg.lstsq(A * ws[:, None], ys * ws, rcond=None)[0]
  yc = (np.cos(r) - 1) / z
  A = np.vander(z, deg_c, increasing=True)
  cc = np.linalg.lstsq(A * z[:, None], yc * z, rcond=None)[0]
  return cs.astype(f32), cc.astype(f32)


def sincos_emul(x, cs, cc, pis, inv_pi):
  x = x.astype(f32)
  magic = f32(12582912.0)
  q = fma(x, np.full_like(x, inv_pi), np.full_like(x, magic))
  k = (q - magic).astype(f32)
  n = q.view(np.int32)
  r = fma(k, np.full_like(x, -pis[0]), x)
  r = fma(k, np.full_like(x, -pis[1]), r)
  r = fma(k, np.full_like(x, -pis[2]), r)
  z = (r * r).astype(f32)
  ps = np.full_like(x, cs[-1])
  for c in cs[-2::-1]:
    ps = fma(ps, z, np.full_like(x, c))
  rz = (r * z).astype(f32)
  s = fma(rz, ps, r)
  pc = np.full_like(x, cc[-1])
  for c in cc[-2::-1]:
    pc = fma(pc, z, np.full_like(x, c))
  c_ = fma(pc, z, np.full_like(x, f32(1.0)))
  sign = np.where(n & 1, f32(-1), f32(1))
  return s * sign, c_ * sign


def main():
  pis = split_pi()
  inv_pi = f32(1 / np.pi)
  print("pi split:", [float.hex(float(p)) for p in pis], [repr(float(p)) for p in pis], "1/pi", repr(float(inv_pi)))
  rng = np.random.default_rng(0)
  for deg_s, deg_c in ((4, 5), (5, 5), (5, 6)):
    cs, cc = fit(deg_s, deg_c, np.pi / 2 * 1.02)
    for span in (8.0, 100.0, 1e4, 1e5):
      x = rng.uniform(-span, span, 2_000_000).astype(f32)
      s, c = sincos_emul(x, cs, cc, pis, inv_pi)
      xs = x.astype(np.float64)
      es, ec = np.abs(s - np.sin(xs)), np.abs(c - np.cos(xs))
      print("deg(s,c)=(%d,%d) |x|<%-8g max abs err sin %.3e cos %.3e   (ulp(1)=1.19e-7)" % (deg_s, deg_c, span, es.max(), ec.max()))
    print("  sin coeffs:", [repr(float(v)) for v in cs])
    print("  cos coeffs:", [repr(float(v)) for v in cc])


if __name__ == "__main__":
  main()
