"""Command line front end with PyCGTOOL's argument surface (``pycgtool/__main__.py:149-254``):

    python -m pycgtool_b200 <topology> [<trajectory>] -m <mapping> -b <bondset> [options]

Orchestration only: load -> ``Mapping.apply`` -> ``BondSet.apply`` -> ``BondSet.boltzmann_invert`` ->
write ``out.gro`` / ``out.itp`` (or an ``.ff`` directory) / ``out.xtc`` / measurement samples.
"""

import argparse
import logging
import pathlib
import sys
import time

from .bondset import BondSet
from .forcefield import ForceField
from .frame import Frame
from .mapping import Mapping

logger = logging.getLogger(__name__)


class ArgumentValidationError(ValueError):
    """Invalid combination of command line arguments."""


class BooleanAction(argparse.Action):
    """``--flag`` / ``--no-flag`` pair."""

    def __init__(self, option_strings, dest, nargs=None, **kwargs):
        super().__init__(option_strings, dest, nargs=0, **kwargs)

    def __call__(self, parser, namespace, values, option_string=None):
        setattr(namespace, self.dest, not option_string.startswith("--no-"))


class PyCGTOOL:
    """One run: the same sequence of steps as the reference driver (``__main__.py:29-112``)."""

    def __init__(self, config):
        self.config = config
        # A mapped .xtc that is used whole is streamed: compressed image up, decoded and mapped block by block on
        # the GPU, the atomistic coordinates never held as a whole (Mapping.apply_xtc)
        streamed = bool(config.mapping and config.trajectory and getattr(config, "gpu_decode", False)
                        and str(config.trajectory).lower().endswith(".xtc") and not config.begin and config.end is None)
        self.in_frame = Frame(topology_file=config.topology, trajectory_file=None if streamed else config.trajectory,
                              frame_start=config.begin, frame_end=config.end,
                              device_decode=getattr(config, "gpu_decode", False))
        self.mapping = None
        self.out_frame = self.in_frame
        if config.mapping:
            self.mapping = Mapping(config.mapping, config, itp_filename=config.itp)
            self.out_frame = (self.mapping.apply_xtc(self.in_frame, config.trajectory) if streamed
                              else self.mapping.apply(self.in_frame))
            self.out_frame.save(self.get_output_filepath("gro"), frame_number=0)
        if self.out_frame.natoms == 0:
            logger.warning("Mapped trajectory contains no atoms - check your mapping file is correct!")
            return
        self.bondset = None
        if config.bondset:
            self.bondset = BondSet(config.bondset, config)
            self.measure_bonds()
        if config.output_xtc:
            self.out_frame.save(self.get_output_filepath("xtc"))

    def get_output_filepath(self, ext):
        out_dir = pathlib.Path(self.config.out_dir)
        out_dir.mkdir(parents=True, exist_ok=True)
        return out_dir.joinpath(self.config.output_name + "." + ext)

    def measure_bonds(self):
        self.bondset.apply(self.out_frame)
        if self.mapping is not None and self.out_frame.n_frames > 1:
            # inversion needs a mapping and more than one frame, else every force constant is infinite
            self.bondset.boltzmann_invert()
            if self.config.output_forcefield:
                forcefield = ForceField(self.config.output_name, dir_path=pathlib.Path(self.config.out_dir),
                                        martini_dir=getattr(self.config, "martini_dir", None))
                forcefield.write(self.config.output_name, self.mapping, self.bondset)
            else:
                self.bondset.write_itp(self.get_output_filepath("itp"), mapping=self.mapping)
        if self.config.dump_measurements:
            self.bondset.dump_values(self.config.dump_n_values, self.config.out_dir)


def parse_arguments(arg_list):
    parser = argparse.ArgumentParser(
        prog="pycgtool_b200", formatter_class=argparse.ArgumentDefaultsHelpFormatter,
        description="Generate coarse-grained molecular dynamics models from atomistic trajectories (B200 GPU).")
    inp = parser.add_argument_group("input files")
    inp.add_argument("topology", type=str, help="AA simulation topology (GRO)")
    inp.add_argument("trajectory", type=str, nargs="?", help="AA simulation trajectory (XTC)")
    inp.add_argument("-m", "--mapping", type=str, help="Mapping file")
    inp.add_argument("-b", "--bondset", type=str, help="Bonds file")
    inp.add_argument("-i", "--itp", type=str, help="GROMACS ITP file")
    inp.add_argument("--begin", type=int, default=0, help="Trajectory frame number to begin")
    inp.add_argument("--end", type=int, default=None, help="Trajectory frame number to end")
    out = parser.add_argument_group("output files")
    out.add_argument("--out-dir", default=".", type=str, help="Directory where output files should be placed")
    out.add_argument("--output-name", default="out", help="Base name of output files")
    out.add_argument("--output-xtc", "--no-output-xtc", default=False, action=BooleanAction,
                     help="Output a pseudo-CG trajectory?")
    out.add_argument("--output", default="gro", help="Coordinate output format")
    out.add_argument("--output-forcefield", "--no-output-forcefield", default=False, action=BooleanAction,
                     help="Output GROMACS forcefield directory?")
    out.add_argument("--martini-dir", default=None, help="Directory with the MARTINI files to copy into the .ff")
    out.add_argument("--dump-measurements", "--no-dump-measurements", default=False, action=BooleanAction,
                     help="Output sample of bond measurements?")
    out.add_argument("--dump-n-values", type=int, default=10000, help="Size of sample of measurements to output")
    mp = parser.add_argument_group("mapping options")
    mp.add_argument("--map-center", default="geom", choices=["geom", "mass", "first"], help="Mapping method")
    mp.add_argument("--virtual-map-center", default="geom", choices=["geom", "mass"],
                    help="Virtual site mapping method")
    bd = parser.add_argument_group("bond options")
    bd.add_argument("--constr-threshold", type=float, default=100000,
                    help="Convert bonds with force constants over [value] to constraints")
    bd.add_argument("--temperature", type=float, default=310, help="Temperature of reference simulation")
    bd.add_argument("--default-fc", "--no-default-fc", default=False, action=BooleanAction,
                    help="Use default MARTINI force constants?")
    bd.add_argument("--generate-angles", "--no-generate-angles", default=True, action=BooleanAction,
                    help="Generate angles from bonds?")
    bd.add_argument("--generate-dihedrals", "--no-generate-dihedrals", default=False, action=BooleanAction,
                    help="Generate dihedrals from bonds?")
    bd.add_argument("--length-form", default="Harmonic", help="Form of bond potential")
    bd.add_argument("--angle-form", default="CosHarmonic", help="Form of angle potential")
    bd.add_argument("--dihedral-form", default="Harmonic", help="Form of dihedral potential")
    run = parser.add_argument_group("run options")
    run.add_argument("--gpu-decode", "--no-gpu-decode", default=True, action=BooleanAction,
                     help="Decode the XTC trajectory on the GPU?")
    run.add_argument("--log-level", default="INFO", choices=("DEBUG", "INFO", "WARNING", "ERROR", "CRITICAL"))
    args = parser.parse_args(arg_list)
    try:
        return validate_arguments(args)
    except ArgumentValidationError as exc:
        parser.error(str(exc))


def validate_arguments(args):
    if not args.dump_measurements:
        args.dump_measurements = bool(args.bondset) and not bool(args.mapping)
    if not args.mapping and not args.bondset:
        raise ArgumentValidationError("One or both of -m and -b is required")
    return args


def main(argv=None):
    start = time.time()
    args = parse_arguments(sys.argv[1:] if argv is None else argv)
    logging.basicConfig(level=args.log_level, format="%(message)s")
    for label, value in (("Topology", args.topology), ("Trajectory", args.trajectory),
                         ("Mapping", args.mapping), ("Bondset", args.bondset)):
        logger.info("%s:\t%s", label, value)
    run = PyCGTOOL(args)
    logger.info("Processed %d frames in %.2f s", run.out_frame.n_frames, time.time() - start)
    return run


if __name__ == "__main__":
    main()
