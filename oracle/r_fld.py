"""ORACLE — TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (a reading of the reference's C++, not a run of it).

The fragment-length prior behind the floating-point columns of fragmentInfo.txt and rawCount.txt, restated
from /root/reference/Circall_v1.0.1/:
  src/Circall_bsj.cpp:832-861     getNormalFragLengthCounts   (mean 200, sd 80, 10 000 samples, 1000 lengths: options :1350,1374-1381)
  src/Circall_bsj.cpp:805-830     getNormalFragLengthDist     (correction factors from the kernel)
  src/Circall_bsj.cpp:977-1006    computeSmoothedEffectiveLengths
  src/EmpiricalDistribution.cpp:36-113,128-146   buildDistribution; median / mean / sd returned as float
"""
import math
import struct


def _f32(x):
    return struct.unpack("f", struct.pack("f", x))[0]


def kernel(p, mean=200, sd=80):
    inv = 1.0 / sd
    x = inv * (p - mean)
    return math.exp(-0.5 * x * x) * inv


def normal_counts(max_len=1000, samples=10000):
    mass = sum(kernel(float(i)) for i in range(max_len))
    return [int(round(kernel(float(i)) * samples / mass)) for i in range(max_len)]      # std::round: halves away from zero; no exact halves here


def correction_factors(max_len=1000):
    out, cm, cd = [0.0] * max_len, 0.0, 0.0
    for i in range(max_len):
        d = kernel(float(i))
        cm += i * d
        cd += d
        if cd > 0:
            out[i] = cm / cd
    return out


def empirical(lens):
    """(median, mean, sd) of EmpiricalDistribution(vals = 0 .. n - 1, lens) as the float accessors return them"""
    n = len(lens)
    valsum = float(sum(lens))
    cumpr, lastval = 0.0, 0
    while lastval < n:
        cumpr += lens[lastval] / valsum
        if cumpr > 1.0 - 1e-6:
            break
        lastval += 1
    upper = lastval + 1
    valsum = float(sum(lens[:upper]))
    pdf = [lens[v] / valsum for v in range(upper)]
    mu = 0.0
    for v in range(1, upper):
        mu += v * pdf[v]
    sd = 0.0
    for i in range(upper):
        sd += (i - mu) * (i - mu) * pdf[i]
    sd = math.sqrt(sd)
    i, j = 0, n - 1
    u, v = lens[0], lens[n - 1]
    while i < j:
        if u <= v:
            v -= u
            i += 1
            u = lens[i]
        else:
            u -= v
            j -= 1
            v = lens[j]
    return _f32(float(i)), _f32(mu), _f32(sd)


def effective_length(ref_len, factors):
    max_len = len(factors)
    cf = factors[max_len - 1] if ref_len >= max_len else factors[ref_len]
    e = float(ref_len) - cf + 1.0
    return float(ref_len) if e < 1.0 else e
