"""Command-line mirror of the reference sub-commands that sit on the replaced path:

    python -m stim_b200 detect --shots N [--in FILE] [--out FILE] [--out_format F] [--seed S]
                               [--append_observables | --prepend_observables] [--obs_out FILE] [--obs_out_format F]
    python -m stim_b200 sample --shots N [--in FILE] [--out FILE] [--out_format F] [--seed S] [--skip_reference_sample]
    python -m stim_b200 sample_dem --shots N [--in FILE] [--out FILE] [--out_format F] [--obs_out FILE] [--obs_out_format F]
                               [--err_out FILE] [--err_out_format F] [--replay_err_in FILE] [--replay_err_in_format F] [--seed S]
    python -m stim_b200 m2d --circuit FILE [--in FILE] [--in_format F] [--out FILE] [--out_format F]
                            [--sweep FILE] [--sweep_format F] [--append_observables] [--obs_out FILE] [--obs_out_format F]
                            [--skip_reference_sample]
    python -m stim_b200 convert --in_format F [--out_format F] [--in FILE] [--out FILE] [--obs_out FILE] [--obs_out_format F]
                            [--num_measurements N] [--num_detectors N] [--num_observables N] [--bits_per_shot N]
                            [--circuit FILE --types MDL] [--dem FILE]

Same flags, defaults and output bytes as `stim detect` / `stim sample` / `stim sample_dem`
(/root/reference/src/stim/cmd/command_detect.cc:23-79, command_sample.cc:25-71, command_sample_dem.cc:25-93, command_m2d.cc,
command_convert.cc:31-244 - a host-only re-encoding of shot data between the formats either side of the path;
every format F is one of 01 b8 r8 hits dets ptb64 on both sides, doc/usage_command_line.md); the sampling itself
runs on the GPU through the C ABI (there is no CPU fallback). Errors print to stderr and exit with status 1 like
/root/reference/src/stim/main_namespaced.cc:113-122."""
import argparse
import sys

import stim_b200

FORMATS = ("01", "b8", "ptb64", "hits", "r8", "dets")


class _Parser(argparse.ArgumentParser):
    """Bad command lines end like the reference's (red message on stderr, exit status 1:
    /root/reference/src/stim/util_bot/arg_parse.cc check_for_unknown_arguments / find_*_argument), not with argparse's status 2."""

    def error(self, message):
        raise ValueError(message)


def _parser():
    p = _Parser(prog="python -m stim_b200", allow_abbrev=False)
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
    for flag in ("out", "obs_out", "err_out", "replay_err_in"):
        q.add_argument("--" + flag, default=None)
        q.add_argument("--" + flag + "_format", default="01", choices=FORMATS)
    q = sub.add_parser("m2d", allow_abbrev=False)
    q.add_argument("--circuit", required=True)
    q.add_argument("--in", dest="inp", default=None)
    q.add_argument("--in_format", default="01", choices=FORMATS)
    q.add_argument("--sweep", default=None)
    q.add_argument("--sweep_format", default="01", choices=FORMATS)
    q.add_argument("--out", default=None)
    q.add_argument("--out_format", default="01", choices=FORMATS)
    q.add_argument("--obs_out", default=None)
    q.add_argument("--obs_out_format", default="01", choices=FORMATS)
    q.add_argument("--append_observables", action="store_true")
    q.add_argument("--skip_reference_sample", action="store_true")
    q = sub.add_parser("convert", allow_abbrev=False)
    q.add_argument("--in_format", required=True, choices=FORMATS)
    q.add_argument("--out_format", default="01", choices=FORMATS)
    q.add_argument("--obs_out_format", default="01", choices=FORMATS)
    q.add_argument("--in", dest="inp", default=None)
    q.add_argument("--out", default=None)
    q.add_argument("--obs_out", default=None)
    q.add_argument("--circuit", default=None)
    q.add_argument("--dem", default=None)
    q.add_argument("--types", default=None)
    for flag in ("num_measurements", "num_detectors", "num_observables", "bits_per_shot"):
        q.add_argument("--" + flag, type=int, default=0)
    return p


def _m2d(args) -> int:
    conv = stim_b200.Circuit(open(args.circuit).read()).compile_m2d_converter(skip_reference_sample=args.skip_reference_sample)
    sys.stdout.flush()
    conv.convert_file(
        measurements_filepath=args.inp if args.inp is not None else "/dev/stdin", measurements_format=args.in_format,
        sweep_bits_filepath=args.sweep, sweep_bits_format=args.sweep_format,
        detection_events_filepath=args.out if args.out is not None else "/dev/stdout", detection_events_format=args.out_format,
        append_observables=args.append_observables, obs_out_filepath=args.obs_out, obs_out_format=args.obs_out_format)
    return 0


def _convert(args) -> int:
    """`stim convert` (/root/reference/src/stim/cmd/command_convert.cc:31-244): the record layout comes from the explicit
    counts, else a detector error model, else a circuit + --types, else --bits_per_shot (anonymous bits; not for dets)."""
    import numpy as np

    from . import _formats

    for flag in ("num_measurements", "num_detectors", "num_observables", "bits_per_shot"):
        if getattr(args, flag) < 0:
            raise ValueError(f"--{flag} must be non-negative.")
    nm, nd, no = args.num_measurements, args.num_detectors, args.num_observables
    inc_m, inc_d, inc_l = nm > 0, nd > 0, no > 0
    if args.dem is not None:
        dem = stim_b200.DetectorErrorModel.from_file(args.dem)
        nd, no = dem.num_detectors, dem.num_observables
        inc_d, inc_l = nd > 0, no > 0
    if args.circuit is not None:
        if args.types is None:
            raise ValueError("--types required when passing circuit")
        circuit = stim_b200.Circuit(open(args.circuit).read())
        nm, nd, no = circuit.num_measurements, circuit.num_detectors, circuit.num_observables
        inc = {"M": inc_m, "D": inc_d, "L": inc_l}
        for c in args.types:
            if c not in inc:
                raise ValueError("Unknown type passed to --types")
            if inc[c]:
                raise ValueError("Each type in types should only be specified once")
            inc[c] = True
        inc_m, inc_d, inc_l = inc["M"], inc["D"], inc["L"]
    if not (inc_m or inc_d or inc_l):
        if args.out_format == "dets":
            raise ValueError("Not enough information given to parse input file to write to dets. Please given a circuit "
                             "with --types, a DEM file, or explicit number of each desired type")
        if args.bits_per_shot == 0:
            raise ValueError("Not enough information given to parse input file.")
        inc_m, nm = True, args.bits_per_shot
    nm, nd, no = (nm if inc_m else 0), (nd if inc_d else 0), (no if inc_l else 0)
    n = nm + nd + no
    data = sys.stdin.buffer.read() if args.inp is None else open(args.inp, "rb").read()
    rows = _formats.read_shots(data, args.in_format, n, num_measurements=nm, num_detectors=nd, num_observables=no)
    shots = rows.shape[0]
    bits = np.unpackbits(rows, axis=1, bitorder="little", count=n) if n else np.zeros((shots, 0), np.uint8)

    def write(lo, hi, segments, path, fmt):
        # segments: (prefix, count) of the value types inside [lo, hi); the dets writer takes up to two of them
        segments = [(p, c) for p, c in segments if c] or [(b"M", 0)]
        if fmt == "dets" and len(segments) == 3:
            with stim_b200._native.open_out(path) as f:
                for r in bits[:, lo:hi]:
                    toks, base = [b"shot"], 0
                    for p, c in segments:
                        toks += [p + str(i).encode() for i in np.flatnonzero(r[base:base + c])]
                        base += c
                    f.write(b" ".join(toks) + b"\n")
            return
        packed = np.packbits(bits[:, lo:hi], axis=1, bitorder="little") if hi > lo else np.zeros((shots, 0), np.uint8)
        stim_b200._write_rows(packed, hi - lo, path, fmt, segments[0][0], segments[-1][0], segments[0][1])

    sys.stdout.flush()
    out_path = args.out if args.out is not None else "/dev/stdout"
    if args.obs_out is not None:
        write(0, nm + nd, [(b"M", nm), (b"D", nd)], out_path, args.out_format)
        write(nm + nd, n, [(b"L", no)], args.obs_out, args.obs_out_format)
    else:
        write(0, n, [(b"M", nm), (b"D", nd), (b"L", no)], out_path, args.out_format)
    return 0


def main(argv=None) -> int:
    try:
        args = _parser().parse_args(argv)
        if args.command == "m2d":
            return _m2d(args)
        if args.command == "convert":
            return _convert(args)
        text = sys.stdin.read() if args.inp is None else open(args.inp).read()
        if args.command == "sample_dem":
            sampler = stim_b200.DetectorErrorModel(text).compile_sampler(seed=args.seed)
            sys.stdout.flush()
            if args.shots > 0:
                sampler.sample_write(
                    args.shots, det_out_file=args.out if args.out is not None else "/dev/stdout", det_out_format=args.out_format,
                    obs_out_file=args.obs_out, obs_out_format=args.obs_out_format, err_out_file=args.err_out,
                    err_out_format=args.err_out_format, replay_err_in_file=args.replay_err_in,
                    replay_err_in_format=args.replay_err_in_format)
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
