"""TEST INFRASTRUCTURE: plain-Python restatement of the reference's velocity samplers with the uniform draws passed in
(ch4/v3/src/Species.cpp:835-869), pinned bit for bit against the compiled reference on CPU (tests/test_oracle_vs_reference.py)
and used to replay the device's Philox streams in the loader / source tests (tests/test_gpu_stochastic.py)."""
import math

K = 1.380648e-23      # all.h:19
PI = 3.141592653      # all.h:20


def sample_vth(r, T, mass):                                    # Species::sampleVth :855-861
    v_th = math.sqrt(2 * K * T / mass)
    comps = []
    for _ in range(3):
        a, b, c = next(r), next(r), next(r)
        comps.append(v_th * (a + b + c - 1.5))
    return math.sqrt(comps[0] * comps[0] + comps[1] * comps[1] + comps[2] * comps[2])


def sample_v3th(r, T, mass):                                   # Species::sampleV3th :862-869
    v_th = sample_vth(r, T, mass)
    theta = 2 * PI * next(r)
    rr = -1.0 + 2 * next(r)
    a = math.sqrt(1 - rr * rr)
    return [v_th * rr, v_th * (math.cos(theta) * a), v_th * (math.sin(theta) * a)]


def sample_reflected(r, v_mag1, n, mass):                      # Species::sampleReflectedVelocity :835-853 (literal)
    v_th = sample_vth(r, 300, mass)
    v_mag2 = v_mag1 + 1.0 * (v_th - v_mag1)
    sin_t = next(r)
    cos_t = math.sqrt(1 - sin_t * sin_t)
    psi = 2 * PI * next(r)
    if n[0] * 1.0 + n[1] * 0.0 + n[2] * 0.0 != 0:
        t1 = [n[1] * 0.0 - n[2] * 0.0, n[2] * 1.0 - n[0] * 0.0, n[0] * 0.0 - n[1] * 1.0]
    else:
        t1 = [n[1] * 0.0 - n[2] * 1.0, n[2] * 0.0 - n[0] * 0.0, n[0] * 1.0 - n[1] * 0.0]
    t2 = [n[1] * t1[2] - n[2] * t1[1], n[2] * t1[0] - n[0] * t1[2], n[0] * t1[1] - n[1] * t1[0]]
    a, b = sin_t * math.cos(psi), math.sin(psi)
    return [v_mag2 * (a * t1[c] + b * t2[c] + cos_t * n[c]) for c in range(3)]
