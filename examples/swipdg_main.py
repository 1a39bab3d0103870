#!/usr/bin/env python
"""The reference's example driver (examples/linearelliptic/swipdg_main.cc:15-87, block-swipdg_main.cc) on the B200 path:

    python examples/swipdg_main.py [--block] [--alu] [directory]

First run: writes ``linearelliptic.swipdg.cfg`` (``linearelliptic.block-swipdg.cfg`` with --block) into the directory and
asks to review it.  Next run: reads it, creates grid / boundary info / problem, initialises the discretization on
cuda:0, solves (once per entry of the ``[parameter]`` section if the problem is parametric) and writes
``<id>.solution[_to_parameter_<n>].vtu``."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from dune_hdd_b200 import discreteproblem  # noqa: E402


def main(argv):
    args = [a for a in argv[1:] if not a.startswith("--")]
    directory = args[0] if args else os.getcwd()
    cls = discreteproblem.LinearellipticExampleBlockSWIPDG if "--block" in argv else discreteproblem.LinearellipticExampleSWIPDG
    info = lambda s: (sys.stdout.write(s), sys.stdout.flush())
    config_file_name = os.path.join(directory, cls.static_id() + ".cfg")
    if not os.path.exists(config_file_name):
        info("Writing default configuration to '%s'... " % config_file_name)
        cls.write_config_file(config_file_name)
        info("done.\nPlease review the configuration and start me again!\n")
        return 0
    example = cls("alu" if "--alu" in argv else "sgrid", device=0, out=info)
    info("initializing discretization... ")
    t = time.perf_counter()
    example.initialize([directory])
    info("done (took %.3gs)\n" % (time.perf_counter() - t))
    discretization = example.discretization()
    prefix = os.path.join(directory, example.static_id())
    if discretization.parametric():
        info("discretization is parametric with parameter_type:\n  %s\n" % discretization.parameter_type())
        parameters = example.discrete_problem().parameters()
        if not parameters:
            info("doing nothing, since there is no 'parameter' specified in the config!\n")
        name = discretization.problem.parameter_name
        for pp, parameter in enumerate(parameters):
            if name not in parameter:
                raise KeyError("parameter %d of the config has no key '%s'" % (pp, name))
            mu = parameter[name]
            info("solving for mu = {%s: %s}... " % (name, mu))
            t = time.perf_counter()
            solution = discretization.solve(mu=mu)
            info(" done (took %.3gs)\n" % (time.perf_counter() - t))
            discretization.visualize(solution, prefix + ".solution_to_parameter_%d" % pp, "solution to parameter %d" % pp)
    else:
        info("discretization is not parametric, solving... ")
        t = time.perf_counter()
        solution = discretization.solve()
        info(" done (took %.3gs)\n" % (time.perf_counter() - t))
        discretization.visualize(solution, prefix + ".solution", "solution")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
