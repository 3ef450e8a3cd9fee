"""Command-line mirror of the reference sub-commands that sit on the replaced path:

    python -m stim_b200 detect --shots N [--in FILE] [--out FILE] [--out_format F] [--seed S]
                               [--append_observables | --prepend_observables] [--obs_out FILE] [--obs_out_format F]
    python -m stim_b200 sample --shots N [--in FILE] [--out FILE] [--out_format F] [--seed S] [--skip_reference_sample]
    python -m stim_b200 sample_dem --shots N [--in FILE] [--out FILE] [--out_format F] [--obs_out FILE] [--obs_out_format F]
                               [--err_out FILE] [--err_out_format F] [--replay_err_in FILE] [--replay_err_in_format F] [--seed S]
    python -m stim_b200 m2d --circuit FILE [--in FILE] [--in_format F] [--out FILE] [--out_format F]
                            [--sweep FILE] [--sweep_format F] [--append_observables] [--obs_out FILE] [--obs_out_format F]
                            [--skip_reference_sample]

Same flags, defaults and output bytes as `stim detect` / `stim sample` / `stim sample_dem`
(/root/reference/src/stim/cmd/command_detect.cc:23-79, command_sample.cc:25-71, command_sample_dem.cc:25-93, command_m2d.cc;
every format F is one of 01 b8 r8 hits dets ptb64 on both sides, doc/usage_command_line.md); the sampling itself
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
