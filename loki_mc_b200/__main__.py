"""python -m loki_mc_b200 SETUP_FILE [NUM_THREADS] -- the reference executable's command line (Sources/lokimc.C) on the GPU engine.
Run from a directory that holds Input/; results go to Output/<output.folder>/.  NUM_THREADS is accepted for compatibility and has no
effect; the number of GPUs a job is sharded over comes from the environment variable LOKIB200_GPUS (default 1)."""
import os
import sys

from . import LokiB200Error, build, lib_path, run_setup


def main(argv):
    if len(argv) not in (2, 3):
        print("usage: [LOKIB200_GPUS=n] python -m loki_mc_b200 SETUP_FILE [NUM_THREADS]")
        return 2
    if not os.path.exists(lib_path()):
        build()
    try:
        os.remove("errorLog.txt")
    except OSError:
        pass
    try:
        run_setup("Input", argv[1], "Output", n_devices=max(1, int(os.environ.get("LOKIB200_GPUS", "1"))))
    except LokiB200Error as e:
        with open("errorLog.txt", "a") as f:
            f.write("Program stopped due to the following error:\n%s\n" % e)
        print("\033[31mProgram stopped due to the following error:\n%s\n\033[0m" % e)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
