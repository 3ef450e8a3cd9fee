"""Command-line mirror of the reference sub-commands that sit on the replaced path:

    python -m stim_b200 detect --shots N [--in FILE] [--out FILE] [--out_format F] [--seed S]
                               [--append_observables | --prepend_observables] [--obs_out FILE] [--obs_out_format F]
    python -m stim_b200 sample --shots N [--in FILE] [--out FILE] [--out_format F] [--seed S] [--skip_reference_sample]
    python -m stim_b200 sample_dem --shots N [--in FILE] [--out FILE] [--out_format F] [--obs_out FILE] [--obs_out_format F]
                               [--err_out FILE] [--err_out_format F] [--seed S]
    python -m stim_b200 m2d --circuit FILE [--in FILE] [--in_format 01|b8] [--out FILE] [--out_format F]
                            [--sweep FILE] [--sweep_format 01|b8] [--append_observables] [--obs_out FILE] [--obs_out_format F]
                            [--skip_reference_sample]

Same flags, defaults and output bytes as `stim detect` / `stim sample` / `stim sample_dem`
(/root/reference/src/stim/cmd/command_detect.cc:23-79, command_sample.cc:25-71, command_sample_dem.cc:25-93, command_m2d.cc;
m2d reads its measurement / sweep data in the 01 and b8 formats only,
doc/usage_command_line.md); the sampling itself
runs on the GPU through the C ABI (there is no CPU fallback). Errors print to stderr and exit with status 1 like
/root/reference/src/stim/main_namespaced.cc:113-122."""
import argparse
import sys

import stim_b200

FORMATS = ("01", "b8", "ptb64", "hits", "r8", "dets")


def _parser():
    p = argparse.ArgumentParser(prog="python -m stim_b200", allow_abbrev=False)
    sub = p.add_subparsers(dest="command", required=True)
    for name in ("detect", "sample"):
        q = sub.add_parser(name, allow_abbrev=False)
        q.add_argument("--shots", type=int, default=1)
        q.add_argument("--in", dest="inp", default=None)
        q.add_argument("--out", default=None)
        q.add_argument("--out_format", default="01", choices=FORMATS)
        q.add_argument("--seed", type=int, default=None)
        if name == "detect":
            q.add_argument("--append_observables", action="store_true")
            q.add_argument("--prepend_observables", action="store_true")
            q.add_argument("--obs_out", default=None)
            q.add_argument("--obs_out_format", default="01", choices=FORMATS)
        else:
            q.add_argument("--skip_reference_sample", action="store_true")
    q = sub.add_parser("sample_dem", allow_abbrev=False)
    q.add_argument("--shots", type=int, default=1)
    q.add_argument("--in", dest="inp", default=None)
    q.add_argument("--seed", type=int, default=None)
    for flag in ("out", "obs_out", "err_out"):
        q.add_argument("--" + flag, default=None)
        q.add_argument("--" + flag + "_format", default="01", choices=FORMATS)
    q = sub.add_parser("m2d", allow_abbrev=False)
    q.add_argument("--circuit", required=True)
    q.add_argument("--in", dest="inp", default=None)
    q.add_argument("--in_format", default="01", choices=("01", "b8"))
    q.add_argument("--sweep", default=None)
    q.add_argument("--sweep_format", default="01", choices=("01", "b8"))
    q.add_argument("--out", default=None)
    q.add_argument("--out_format", default="01", choices=FORMATS)
    q.add_argument("--obs_out", default=None)
    q.add_argument("--obs_out_format", default="01", choices=FORMATS)
    q.add_argument("--append_observables", action="store_true")
    q.add_argument("--skip_reference_sample", action="store_true")
    return p


def _read_rows(data: bytes, fmt: str, n_bits: int):
    """01 / b8 records -> packed uint8 rows (measure_record_reader.inl: one record per shot)."""
    import numpy as np

    nb = (n_bits + 7) // 8
    if fmt == "b8":
        if nb == 0:
            raise ValueError("b8 data with zero bits per shot does not say how many shots there are.")
        if len(data) % nb:
            raise ValueError("b8 data ended in the middle of a record.")
        return np.frombuffer(data, dtype=np.uint8).reshape(-1, nb).copy()
    lines = data.decode().split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    bits = np.zeros((len(lines), n_bits), dtype=np.uint8)
    for i, ln in enumerate(lines):
        if len(ln) != n_bits or ln.strip("01"):
            raise ValueError("01 data didn't have the expected number of 0/1 characters per line.")
        bits[i] = np.frombuffer(ln.encode(), dtype=np.uint8) - 48
    return np.packbits(bits, axis=1, bitorder="little") if n_bits else np.zeros((len(lines), 0), dtype=np.uint8)


def _m2d(args) -> int:
    import ctypes

    from . import _native

    conv = stim_b200.Circuit(open(args.circuit).read()).compile_m2d_converter(skip_reference_sample=args.skip_reference_sample)
    data = sys.stdin.buffer.read() if args.inp is None else open(args.inp, "rb").read()
    meas = _read_rows(data, args.in_format, conv.num_measurements)
    sweep = None if args.sweep is None else _read_rows(open(args.sweep, "rb").read(), args.sweep_format, conv.num_sweep_bits)
    D, L = conv.num_detectors, conv.num_observables
    want_obs_file = args.obs_out is not None
    res = conv.convert(measurements=meas, sweep_bits=sweep, append_observables=args.append_observables,
                       separate_observables=want_obs_file, bit_packed=True)
    dets, obs = res if want_obs_file else (res, None)

    def write(rows, n_bits, path, fmt, p1, p2, transition):
        rows = rows if rows.size else rows.reshape(rows.shape[0], 0)
        with open(path, "wb") as f:
            buf = rows.ctypes.data_as(ctypes.c_void_p) if rows.size else None
            _native.check(_native.lib().gstim_write_shots_to_fd(
                buf, rows.shape[1], rows.shape[0], n_bits, f.fileno(), fmt.encode(), p1, p2, transition))

    sys.stdout.flush()
    write(dets, D + (L if args.append_observables else 0), args.out if args.out is not None else "/dev/stdout", args.out_format,
          b"D", b"L", D)
    if want_obs_file:
        write(obs, L, args.obs_out, args.obs_out_format, b"L", b"L", L)
    return 0


def main(argv=None) -> int:
    args = _parser().parse_args(argv)
    try:
        if args.command == "m2d":
            return _m2d(args)
        text = sys.stdin.read() if args.inp is None else open(args.inp).read()
        if args.command == "sample_dem":
            sampler = stim_b200.DetectorErrorModel(text).compile_sampler(seed=args.seed)
            sys.stdout.flush()
            if args.shots > 0:
                sampler.sample_write(
                    args.shots, det_out_file=args.out if args.out is not None else "/dev/stdout", det_out_format=args.out_format,
                    obs_out_file=args.obs_out, obs_out_format=args.obs_out_format, err_out_file=args.err_out,
                    err_out_format=args.err_out_format)
            return 0
        circuit = stim_b200.Circuit(text)
        out_path = args.out if args.out is not None else "/dev/stdout"
        sys.stdout.flush()
        if args.command == "detect":
            if args.prepend_observables and args.append_observables:
                raise ValueError("--prepend_observables and --append_observables are mutually exclusive.")
            sampler = circuit.compile_detector_sampler(seed=args.seed)
            if args.shots > 0:
                sampler.sample_write(
                    args.shots, filepath=out_path, format=args.out_format, obs_out_filepath=args.obs_out,
                    obs_out_format=args.obs_out_format, prepend_observables=args.prepend_observables,
                    append_observables=args.append_observables)
        else:
            sampler = circuit.compile_sampler(skip_reference_sample=args.skip_reference_sample, seed=args.seed)
            if args.shots > 0:
                sampler.sample_write(args.shots, filepath=out_path, format=args.out_format)
        return 0
    except (ValueError, IndexError, OSError, RuntimeError) as ex:
        sys.stderr.write("\033[31m" + str(ex) + "\033[0m\n")
        return 1


if __name__ == "__main__":
    sys.exit(main())
