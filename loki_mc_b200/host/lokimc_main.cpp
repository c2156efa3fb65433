// lokimc_main.cpp -- command-line front end with the reference executable's conventions (Sources/lokimc.C:45-95, Sources/Message.C):
//   lokimc_b200 SETUP_FILE [NUM_GPUS]
// run from a directory that holds Input/ (setup, LXCat and database files); results go to Output/<output.folder>/; errors are
// appended to errorLog.txt and end the program with a non-zero status.  NUM_GPUS takes the place of the reference's NUM_THREADS.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../include/lokib200_host.h"

int main(int argc, char* argv[]) {
  std::remove("errorLog.txt");
  std::string setupFile;
  int nGpus = 1;
  if (argc == 2) setupFile = argv[1];
  else if (argc == 3) { setupFile = argv[1]; nGpus = std::atoi(argv[2]); }
  else {
    std::printf("Insert the name of the setup file and the number of GPUs in the following form:\nSETUP_FILE  NUM_GPUS\n");
    char name[512];
    if (std::scanf("%511s %d", name, &nGpus) != 2) return EXIT_FAILURE;
    setupFile = name;
  }
  if (nGpus < 1) nGpus = 1;
  const int available = lokib200_device_count();
  if (available >= 1 && nGpus > available) nGpus = available;
  if (lokib200_run_setup("Input", setupFile.c_str(), "Output", nGpus, 0, 1, nullptr) != 0) {
    const char* msg = lokib200_run_last_error();
    if (FILE* f = std::fopen("errorLog.txt", "a")) { std::fprintf(f, "Program stopped due to the following error:\n%s\n", msg); std::fclose(f); }
    std::printf("\033[31mProgram stopped due to the following error:\n%s\n\033[0m", msg);
    return EXIT_FAILURE;
  }
  return 0;
}
