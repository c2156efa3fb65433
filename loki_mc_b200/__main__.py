"""python -m loki_mc_b200 SETUP_FILE [NUM_GPUS] -- the reference executable's command line (Sources/lokimc.C) on the GPU engine.
Run from a directory that holds Input/; results go to Output/<output.folder>/."""
import os
import sys

from . import LokiB200Error, build, lib_path, run_setup


def main(argv):
    if len(argv) not in (2, 3):
        print("usage: python -m loki_mc_b200 SETUP_FILE [NUM_GPUS]")
        return 2
    if not os.path.exists(lib_path()):
        build()
    try:
        os.remove("errorLog.txt")
    except OSError:
        pass
    try:
        run_setup("Input", argv[1], "Output", n_devices=int(argv[2]) if len(argv) == 3 else 1)
    except LokiB200Error as e:
        with open("errorLog.txt", "a") as f:
            f.write("Program stopped due to the following error:\n%s\n" % e)
        print("\033[31mProgram stopped due to the following error:\n%s\n\033[0m" % e)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
