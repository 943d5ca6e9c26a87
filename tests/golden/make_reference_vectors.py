"""Writes tests/golden/reference_vectors.json: every known-answer vector the reference's own unit tests hold for
the prove/verify path, transcribed by hand from the cited lines (the Rust crate cannot be run here: no cargo).
The literals below ARE the fixture; the script only serialises them so the JSON can be regenerated and diffed.
Citations are relative to the reference repository.  Run: python tests/golden/make_reference_vectors.py
"""
import json
import os

V = {
    # src/utils/u64field.rs:238-254 (F_101)
    "field_f101": [
        {"op": "add", "a": 100, "b": 100, "out": 200 % 101},
        {"op": "sub", "a": 0, "b": 1, "out": 100},
        {"op": "div", "a": 1, "b": 0, "out": None},
        {"op": "div_then_mul", "a": 4, "b": 12, "out": 4},            # 12 * (4/12) == 4
        {"op": "neg", "a": 1, "out": 100},
        {"op": "neg_div", "a": 1, "b": 2, "out": 50},
        {"op": "neg_div", "a": 1, "b": 5, "out": 20},
        {"op": "pow", "a": 100, "b": 0, "out": 1},
        {"op": "pow", "a": 100, "b": 2, "out": (100 * 100) % 101},
        {"op": "pow", "a": 100, "b": 3, "out": (100 * 100 * 100) % 101},
    ],
    # src/poly.rs:402-487 (F_15485863)
    "poly_f15485863": {
        "modulus": 15485863,
        "add": [[[1, 2, 3], [1, 2, 3], [2, 4, 6]], [[1, 2, 3], [1, 2, 3, 4, 5], [2, 4, 6, 4, 5]], [[1, 2, 3, 4, 6], [1, 2, 3], [2, 4, 6, 4, 6]]],
        "sub": [[[1, 2, 3], [1, 2, 3], [0]], [[1, 2, 3], [1, 2], [0, 0, 3]]],
        "mul": [[[5, 0, 10, 6], [1, 2, 4], [5, 10, 30, 26, 52, 24]]],
        "div_roundtrip": [[[1], [1, 1]], [[1, 1], [1, 1]], [[1, 2, 1], [1, 1]], [[1, 2, 1, 2, 5, 8, 1, 9], [1, 1, 5, 4]]],
        "lagrange_points": [[1, 2], [5, 7], [7, 9], [3, 1]],
        "z": [[[1, 5], [5, -6, 1]]],
        "eval": [[[1, 2, 1], 2, 9]],
        "normalize": [[[1, 0, 0, 0], [1]]],
    },
    # src/matrix.rs:203-227 (F_104729)
    "matrix_f104729": {
        "modulus": 104729,
        "add": {"a": [1, 2], "b": [3, 4], "shape": [2, 1], "out": [4, 6]},
        "mul": {"a": [1, 2, 3, 4, 5, 6], "ashape": [2, 3], "b": [10, 11, 20, 21, 30, 31], "bshape": [3, 2], "out": [140, 146, 320, 335]},
        "inv_involution": {"a": [1, 2, 3, 4, 1, 6, 7, 8, 9], "shape": [3, 3]},
    },
    # src/fft.rs:140-183 (F_337, omega = 85, n = 8)
    "fft_f337": {
        "modulus": 337, "omega": 85, "size": 8,
        "values": [3, 1, 4, 1, 5, 9, 2, 6], "freq": [31, 70, 109, 74, 334, 181, 232, 4],
        "mul_a": [24, 12, 28, 8], "mul_b": [4, 26, 29, 23],
    },
    # src/pbh/g1.rs:233-260
    "g1": {
        "generator": [1, 2],
        "neg_g": [1, 99], "two_g": [68, 74], "neg_two_g": [68, 27], "four_g": [65, 98], "neg_four_g": [65, 3],
        "eight_g": [18, 49], "neg_eight_g": [18, 52], "sixteen_g": [1, 99], "neg_sixteen_g": [1, 2],
        "two_g_plus_g": [26, 45], "four_g_plus_g": [12, 32], "eight_g_plus_g": [18, 52],
    },
    # src/pbh/g2.rs:108-119
    "g2": {"generator": [36, 31], "two_g": [90, 82]},
    # src/pbh/gt.rs:88-97
    "gt": {
        "mul": [[[26, 97], [93, 76], [97, 89]]],
        "pow": [[[42, 49], 6, [97, 89]], [[68, 47], 600, [97, 89]]],
        "frobenius_base": [93, 76],
    },
    # src/pbh/pairing.rs:56-75: P = G, R = 4G, Q = 3*G2, a = 5 (equalities only)
    "pairing": {"p_scalar": 1, "r_scalar": 4, "q_scalar": 3, "a": 5},
    # src/pbh/mod.rs:44-124 — the only end-to-end vector; README.md:6 names the pairing value
    "plonk_gen_proof": {
        "s": 2, "srs_n": 6, "omega_pows": 4,
        "a": [3, 4, 5, 9], "b": [3, 4, 5, 16], "c": [9, 16, 25, 25],
        "rand": [7, 4, 11, 12, 16, 2, 14, 11, 7],
        "challange": {"alpha": 15, "beta": 12, "gamma": 13, "z": 5, "v": 12},
        "proof": {"a_s": [91, 66], "b_s": [26, 45], "c_s": [91, 35], "z_s": [32, 59], "t_lo_s": [12, 32], "t_mid_s": [26, 45],
                  "t_hi_s": [91, 66], "w_z_s": [91, 35], "w_z_omega_s": [65, 98], "a_z": 15, "b_z": 13, "c_z": 5,
                  "s_sigma_1_z": 1, "s_sigma_2_z": 12, "r_z": 15, "z_omega_z": 15},
        "verify_rand": [4], "verify": True, "pairing_value": [93, 76],
    },
}

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")
    with open(out, "w") as f:
        json.dump(V, f, indent=1, sort_keys=True)
    print("wrote", out)
