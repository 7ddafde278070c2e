#!/usr/bin/env python
"""Generate tests/golden/*.json: the known-answer vectors the REFERENCE's own tests hold
for the hot path, transcribed from /root/reference/tests and evaluated here with plain
Python floats (math module only -- independent of oracle/ and of lerax_b200/).

The reference (JAX) cannot be imported in this image, so the expected values are those
the reference tests themselves state (closed forms / the in-test loop), not outputs of
a reference run.  Sources:
  gae_*                /root/reference/tests/advantage/test_gae.py:8-19 (loop), :22-47, :50-61, :64-75, :78-85
  rollout_buffer_gae   /root/reference/tests/test_buffer.py:61-82
  categorical_*        /root/reference/tests/distribution/test_categorical.py:49-69, :82-96, :106-112
  normal_*             /root/reference/tests/distribution/test_normal.py:49-65, :75-82
  philox_kat           Random123 kat_vectors (Philox4x32-10), Salmon et al. SC'11

Run:  python tests/golden/make_golden.py     (rewrites the json files in place)
"""

import json
import math
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_gae(rewards, values, dones, last_value, gamma, lam):
    # tests/advantage/test_gae.py:8-19, in float64
    T = len(rewards)
    adv = [0.0] * T
    nxt = 0.0
    for i in range(T - 1, -1, -1):
        nv = last_value if i == T - 1 else values[i + 1]
        nd = 1.0 - float(dones[i])
        delta = rewards[i] + gamma * nv * nd - values[i]
        nxt = delta + gamma * lam * nd * nxt
        adv[i] = nxt
    return adv, [a + v for a, v in zip(adv, values)]


def gae_cases():
    cases = []

    def add(name, r, v, d, last, gamma, lam, expect_adv=None, expect_ret=None, src=""):
        adv, ret = reference_gae(r, v, d, last, gamma, lam)
        if expect_adv is not None:
            assert all(abs(a - b) < 1e-12 for a, b in zip(adv, expect_adv)), name
        if expect_ret is not None:
            assert all(abs(a - b) < 1e-12 for a, b in zip(ret, expect_ret)), name
        cases.append(dict(name=name, src=src, rewards=r, values=v, dones=[bool(x) for x in d],
                          last_value=last, gamma=gamma, lam=lam, advantages=adv, returns=ret))

    add("matches_reference", [1.0, -0.5, 2.0, 0.25], [0.3, 0.1, -0.2, 0.4], [0, 1, 0, 0], 0.75, 0.9, 0.8,
        src="test_gae.py:22-47")
    r, v, d, last, g = [1.0, 2.0, 3.0], [0.5, 0.25, 1.0], [0, 0, 1], 4.0, 0.9
    nv = [0.25, 1.0, 4.0]
    td = [r[i] + g * nv[i] * (1 - d[i]) - v[i] for i in range(3)]
    add("lambda_zero_is_one_step_td", r, v, d, last, g, 0.0, expect_adv=td, src="test_gae.py:50-61")
    add("done_cuts_propagation", [1.0] * 4, [0.0] * 4, [0, 1, 0, 0], 0.0, 1.0, 1.0,
        expect_adv=[2.0, 1.0, 2.0, 1.0], expect_ret=[2.0, 1.0, 2.0, 1.0], src="test_gae.py:64-75")
    add("lambda_one_bootstraps", [0.0] * 3, [0.0] * 3, [0, 0, 0], 3.0, 1.0, 1.0,
        expect_ret=[3.0, 3.0, 3.0], src="test_gae.py:78-85")
    add("jittable_shapes", [1.0] * 8, [0.0] * 8, [0] * 8, 0.0, 0.99, 0.95, src="test_gae.py:88-101")
    add("rollout_buffer_gae", [1.0, 1.0], [0.5, 0.25], [0, 1], 0.0, 0.99, 0.95, src="test_buffer.py:61-82")
    return cases


def distribution_cases():
    ln = math.log
    cat = [
        dict(name="log_prob", src="test_categorical.py:60-69", probs=[0.25, 0.25, 0.5],
             actions=[0, 2], log_probs=[ln(0.25), ln(0.5)]),
        dict(name="entropy_uniform4", src="test_categorical.py:82-88", probs=[0.25] * 4, entropy=ln(4.0)),
        dict(name="entropy_deterministic", src="test_categorical.py:90-96", probs=[1.0, 0.0, 0.0], entropy=0.0),
        dict(name="mode", src="test_categorical.py:106-112", probs=[0.1, 0.2, 0.7], mode=2),
    ]
    nor = [
        dict(name="log_prob_at_mean", src="test_normal.py:49-55", loc=0.0, scale=1.0, x=0.0,
             log_prob=-0.5 * ln(2 * math.pi)),
        dict(name="log_prob_at_one", src="test_normal.py:57-65", loc=0.0, scale=1.0, x=1.0,
             log_prob=-0.5 * ln(2 * math.pi) - 0.5),
        dict(name="entropy_scale2", src="test_normal.py:75-82", loc=0.0, scale=2.0,
             entropy=0.5 * ln(2 * math.pi * math.e * 4.0)),
    ]
    return dict(categorical=cat, normal=nor)


def philox_cases():
    # Random123 kat_vectors, "philox4x32 10"
    return [
        dict(counter=[0, 0, 0, 0], key=[0, 0],
             out=[0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        dict(counter=[0xFFFFFFFF] * 4, key=[0xFFFFFFFF] * 2,
             out=[0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        dict(counter=[0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], key=[0xA4093822, 0x299F31D0],
             out=[0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]


def main():
    for name, obj in (("gae.json", gae_cases()), ("distributions.json", distribution_cases()),
                      ("philox.json", philox_cases())):
        with open(os.path.join(HERE, name), "w") as f:
            json.dump(obj, f, indent=1)
        print("wrote", name)


if __name__ == "__main__":
    main()
