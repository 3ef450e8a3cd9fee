"""Builds the circuit variants of BASELINE.json configs 4 and 5 that the reference's generator cannot emit
 and writes them next to the generated fixtures in tests/golden/circuits/:

  c4v_color_d15_r15_mpp_dense.stim   color code d=15 r=15 where every round's CX + MR ancilla cycle is replaced by one
                                     MPP(p) of each plaquette's Z product (same record layout, so the generator's
                                     detectors stay valid) and dense noise layers are added on the data qubits
                                     (PAULI_CHANNEL_1, DEPOLARIZE1(0.05), a correlated E / ELSE pair)
  c4v_color_d15_r15_mpp_det.stim     the same circuit with every probability in {0, 1} and only deterministic channels
                                     (X_ERROR / PAULI_CHANNEL_1 with a single term / MPP(1) on a subset): reference output
                                     is one fixed row for every shot -> byte-exact ptb64 / b8 comparison
  c5_feedback()                      (not written: 1 MB; tests and tools/gen_stats_big.py call it) surface code d=51 r=51 with
                                     classical feedback after every MR layer (CX / CZ / CY rec[-k] q), for compile_sampler

Pure text transforms of the committed fixtures; needs no reference build."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CDIR = os.path.join(ROOT, "tests", "golden", "circuits")


def _read(name):
    with open(os.path.join(CDIR, name)) as f:
        return f.read()


def _targets(line):
    return [int(t) for t in line.split()[1:]]


def c4_variant(deterministic: bool) -> str:
    src = _read("c4_color_d15_r15.stim").split("\n")
    out = []
    i = 0
    n_rounds = 0
    while i < len(src):
        ln = src[i]
        s = ln.strip()
        indent = ln[: len(ln) - len(ln.lstrip())]
        if s.startswith("CX "):
            # a run of "CX ... / DEPOLARIZE2 ... / TICK" layers followed by "X_ERROR / MR / X_ERROR": collect plaquettes
            nbrs = {}
            j = i
            while src[j].strip().startswith("CX "):
                t = _targets(src[j])
                for k in range(0, len(t), 2):
                    nbrs.setdefault(t[k + 1], []).append(t[k])
                j += 1
                assert src[j].strip().startswith("DEPOLARIZE2"), src[j][:40]
                j += 1
                assert src[j].strip() == "TICK"
                j += 1
            assert src[j].strip().startswith("X_ERROR"), src[j][:40]
            assert src[j + 1].strip().startswith("MR "), src[j + 1][:40]
            assert src[j + 2].strip().startswith("X_ERROR"), src[j + 2][:40]
            anc = _targets(src[j + 1])
            data = sorted({d for a in anc for d in nbrs[a]})
            prods = " ".join("*".join(f"Z{d}" for d in sorted(nbrs[a])) for a in anc)
            if deterministic:
                # deterministic channels only; flips on a rotating subset so detectors differ from round to round
                sub = data[n_rounds % 3::3]
                out.append(f"{indent}X_ERROR(1) " + " ".join(map(str, sub)))
                out.append(f"{indent}PAULI_CHANNEL_1(0, 1, 0) " + " ".join(map(str, data[(n_rounds + 1) % 5::5])))
                out.append(f"{indent}MPP({1 if n_rounds % 2 else 0}) {prods}")
            else:
                out.append(f"{indent}PAULI_CHANNEL_1(0.02, 0.01, 0.03) " + " ".join(map(str, data)))
                out.append(f"{indent}DEPOLARIZE1(0.05) " + " ".join(map(str, data)))
                out.append(f"{indent}E(0.04) X{data[0]} Y{data[1]} Z{data[2]}")
                out.append(f"{indent}ELSE_CORRELATED_ERROR(0.25) Z{data[3]} X{data[4]}")
                out.append(f"{indent}MPP(0.01) {prods}")
            out.append(f"{indent}TICK")
            n_rounds += 1
            i = j + 3
            continue
        if deterministic:
            ln = re.sub(r"\(0\.001\)", "(0)", ln)
        out.append(ln)
        i += 1
    assert n_rounds == 2  # the two REPEAT bodies (text blocks, not executed rounds)
    return "\n".join(out)


def c5_feedback() -> str:
    src = _read("c5_surface_x_d51_r51.stim").split("\n")
    out = []
    for ln in src:
        out.append(ln)
        s = ln.strip()
        if s.startswith("MR "):
            indent = ln[: len(ln) - len(ln.lstrip())]
            t = _targets(ln)
            # feedback from this layer's results into data qubits (odd indices are data qubits in this layout)
            out.append(f"{indent}CX rec[-1] 1 rec[-3] 5 rec[-{len(t)}] 9")
            out.append(f"{indent}CZ rec[-2] 3 7 rec[-5]")
            out.append(f"{indent}CY rec[-7] 11")
    return "\n".join(out)


def main():
    files = {
        "c4v_color_d15_r15_mpp_dense.stim": c4_variant(False),
        "c4v_color_d15_r15_mpp_det.stim": c4_variant(True),
    }
    for name, text in files.items():
        with open(os.path.join(CDIR, name), "w") as f:
            f.write(text)
        print(name, len(text), "bytes")


if __name__ == "__main__":
    main()
