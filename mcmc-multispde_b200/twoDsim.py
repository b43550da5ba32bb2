#!/usr/bin/env python
"""Driver with the flag surface of /root/reference/twoDsim.py (argparse block 68-97) on top of mspde.image.

``--distributed`` (not in the reference) runs ONE chain block-distributed over the ranks of a torchrun launch (n=96).
FBP initialisation (seq-no 0) and chain continuation from the previous sequence's result file (seq-no > 0) follow
twoDsim.py:133-136 through mspde.io.  Not carried over (out of the hot-path scope): Slurm self-resubmission and
post_analysis plotting."""
import argparse
import datetime
import os
import pathlib
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from mspde import image as im, io      # noqa: E402


class _TriStateFlag(argparse.Action):
    """Boolean option in the reference CLI's convention (twoDsim.py:91-96): ``--flag`` and ``--flag true|false`` set it,
    ``--noflag`` clears it.  One action object serves both spellings."""

    def __init__(self, option_strings, dest, default=False, help=None):
        super().__init__(option_strings, dest, nargs="?", default=default, help=help)

    def __call__(self, parser, namespace, values, option_string=None):
        seen = namespace.__dict__.setdefault("_flag_spellings", {})
        negated = option_string.startswith("--no")
        if seen.setdefault(self.dest, negated) != negated:      # the two spellings exclude each other, as in the reference
            parser.error(f"argument {option_string}: not allowed with the opposite spelling of the same flag")
        if negated:
            if values is not None:
                parser.error(f"{option_string} takes no value")
            setattr(namespace, self.dest, False)
            return
        text = "true" if values is None else str(values).strip().lower()
        if text not in ("true", "false"):
            parser.error(f"{option_string} expects true or false, got {values!r}")
        setattr(namespace, self.dest, text == "true")


def add_boolean_argument(parser, name, default=False, messages=""):
    parser.add_argument("--" + name, "--no" + name, dest=name.replace("-", "_"), action=_TriStateFlag, default=default, help=messages)


def build_parser():
    """The reference driver's option schema (twoDsim.py:68-95): same names, types and defaults."""
    parser = argparse.ArgumentParser()
    parser.add_argument('--iter', default=0, type=int)
    parser.add_argument('--seq-no', default=0, type=int)
    parser.add_argument('--n-layers', default=2, type=int)
    parser.add_argument('--n-theta', default=45, type=int)
    parser.add_argument('--n', '--basis-n', dest='n', default=32, type=int)   # --basis-n: alias that survives torchrun's option parser
    parser.add_argument('--seed', default=1, type=int)
    parser.add_argument('--n-extended', default=64, type=int)
    parser.add_argument('--n-samples', default=100, type=int)
    parser.add_argument('--evaluation-interval', default=10, type=int)
    parser.add_argument('--beta', default=1, type=float)
    parser.add_argument('--kappa', default=5e9, type=float)
    parser.add_argument('--chol-epsilon', default=1e-6, type=float)
    parser.add_argument('--sigma-0', default=4e7, type=float)
    parser.add_argument('--sigma-v', default=3.2e4, type=float)
    parser.add_argument('--meas-std', default=0.2, type=float)
    parser.add_argument('--sigma-scaling', default=4e-1, type=float)
    parser.add_argument('--burn-percentage', default=25.0, type=float)
    parser.add_argument('--variant', default="dunlop", type=str)
    parser.add_argument('--phantom-name', default="shepp.png", type=str)
    parser.add_argument('--meas-type', default="tomo", type=str)
    parser.add_argument('--init-folder', default="", type=str)
    parser.add_argument('--result-dir', default="", type=str, help='where result_<seq>.npz is written (default ./SimulationResult)')
    add_boolean_argument(parser, 'naive', default=False)
    add_boolean_argument(parser, 'enable-step-adaptation', default=True)
    add_boolean_argument(parser, 'print-progress', default=True)
    add_boolean_argument(parser, 'verbose', default=True)
    add_boolean_argument(parser, 'hybrid', default=False)
    add_boolean_argument(parser, 'adapt-sqrtbeta', default=False)
    # not in the reference: ONE chain block-distributed over the ranks of a torchrun launch (BASELINE config 5, mspde.dist)
    add_boolean_argument(parser, 'distributed', default=False)
    return parser


def parse_args(argv=None):
    args = build_parser().parse_args(argv)
    args.__dict__.pop("_flag_spellings", None)
    if args.n_theta == 18:                      # twoDsim.py:98-99
        args.seed = 2
    return args


def run_distributed(args):
    """--distributed under torchrun: every rank steps the same chain collectively (mspde.dist.DistSimulation); rank 0 saves."""
    import torch
    import torch.distributed as dist
    from mspde.dist import DistSimulation
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if args.adapt_sqrtbeta or args.naive or args.seq_no != 0:
        raise SystemExit("--distributed covers the plain pCN step from an FBP-free prior start (no --adapt-sqrtbeta, --naive, --seq-no)")
    sim = DistSimulation(n_layers=args.n_layers, n_samples=args.n_samples, n=args.n, n_extended=args.n_extended, beta=args.beta,
                         kappa=args.kappa, sigma_0=args.sigma_0, sigma_v=args.sigma_v, sigma_scaling=args.sigma_scaling,
                         meas_std=args.meas_std, evaluation_interval=args.evaluation_interval,
                         printProgress=args.print_progress and int(os.environ.get("RANK", "0")) == 0, seed=args.seed,
                         burn_percentage=args.burn_percentage, enable_step_adaptation=args.enable_step_adaptation,
                         pcn_variant=args.variant, phantom_name=args.phantom_name, meas_type=args.meas_type, n_theta=args.n_theta,
                         verbose=args.verbose, hybrid_GPU_CPU=args.hybrid, device=local_rank, chol_epsilon=args.chol_epsilon)
    sim.pcn.target_acceptance_rate = 0.234                      # twoDsim.py:107
    sim.run()
    parent = pathlib.Path(args.result_dir) if args.result_dir else pathlib.Path.cwd() / 'SimulationResult'
    folder = parent / ('result-' + datetime.datetime.now().strftime('%d-%b-%Y_%H_%M_%S'))
    if int(os.environ.get("RANK", "0")) == 0:
        folder.mkdir(parents=True, exist_ok=True)
        out = folder / 'result_0.npz'
        sim.save(str(out))
        print('simulation result stored in {} ({:.3f} it/s, acceptance {:.1%})'.format(
            out, args.n_samples / max(sim.total_time, 1e-9), sim.acceptancePercentage))
    if dist.is_initialized():
        dist.destroy_process_group()


def main(argv=None):
    args = parse_args(argv)
    if args.distributed:
        return run_distributed(args)
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    sim = im.Simulation(n_layers=args.n_layers, n_samples=args.n_samples, n=args.n, n_extended=args.n_extended, beta=args.beta,
                        kappa=args.kappa, sigma_0=args.sigma_0, sigma_v=args.sigma_v, sigma_scaling=args.sigma_scaling,
                        meas_std=args.meas_std, evaluation_interval=args.evaluation_interval, printProgress=args.print_progress,
                        seed=args.seed + int(os.environ.get("RANK", "0")), burn_percentage=args.burn_percentage,
                        enable_step_adaptation=args.enable_step_adaptation, pcn_variant=args.variant,
                        phantom_name=args.phantom_name, meas_type=args.meas_type, n_theta=args.n_theta, verbose=args.verbose,
                        hybrid_GPU_CPU=args.hybrid, device=local_rank)
    sim.pcn.set_chol_epsilon(args.chol_epsilon)                 # twoDsim.py:106
    sim.pcn.target_acceptance_rate = 0.234                      # twoDsim.py:107
    sim.pcn.use_beta_adaptation = args.adapt_sqrtbeta           # twoDsim.py:108
    sim.pcn.use_naive_logRatio_implementation = args.naive      # twoDsim.py:109
    parent = pathlib.Path(args.result_dir) if args.result_dir else pathlib.Path.cwd() / 'SimulationResult'
    folder = parent / ('result-' + datetime.datetime.now().strftime('%d-%b-%Y_%H_%M_%S') if args.seq_no == 0 else args.init_folder)
    folder.mkdir(parents=True, exist_ok=True)
    rank_tag = os.environ.get("RANK", "0")
    if args.seq_no == 0:                                         # twoDsim.py:133-136
        io.initialize_using_FBP(sim)
    else:
        prev = folder / 'result_{}_rank{}.npz'.format(args.seq_no - 1, rank_tag)
        if prev.exists():
            io.initialize_from_file(sim, str(prev))              # initialize_from_folder, 36-54
        else:
            io.initialize_using_FBP(sim)                         # the reference's fallback when the file is missing (43-44)
    sim.run()
    out = folder / 'result_{}_rank{}.npz'.format(args.seq_no, rank_tag)
    sim.save(str(out))
    print('simulation result stored in {} ({:.2f} it/s, acceptance {:.1%})'.format(
        out, args.n_samples / max(sim.total_time, 1e-9), sim.acceptancePercentage))


if __name__ == '__main__':
    main()
