// lokimc_main.cpp -- command-line front end with the reference executable's conventions (Sources/lokimc.C:45-95, Sources/Message.C):
//   lokimc_b200 SETUP_FILE [NUM_THREADS]
// run from a directory that holds Input/ (setup, LXCat and database files); results go to Output/<output.folder>/; errors are
// appended to errorLog.txt and end the program with a non-zero status.  The second argument keeps the reference's meaning (OpenMP
// threads, lokimc.C:66-78) and is accepted for compatibility: the ensemble runs on the GPU, so it has no effect.  The number of GPUs a
// job is sharded over comes from the environment variable LOKIB200_GPUS (default 1; "all" = every visible device).
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../include/lokib200_host.h"

int main(int argc, char* argv[]) {
  std::remove("errorLog.txt");
  std::string setupFile;
  int nThreads = 1;
  if (argc == 2) setupFile = argv[1];
  else if (argc == 3) { setupFile = argv[1]; nThreads = std::atoi(argv[2]); }
  else {
    std::printf("Insert the name of the setup file and the number of threads in the following form:\nSETUP_FILE  NUM_THREADS\n");   // lokimc.C:79-84
    char name[512];
    if (std::scanf("%511s %d", name, &nThreads) != 2) return EXIT_FAILURE;
    setupFile = name;
  }
  (void)nThreads;
  const int available = lokib200_device_count();
  int nGpus = 1;
  if (const char* env = std::getenv("LOKIB200_GPUS")) nGpus = (std::string(env) == "all") ? available : std::atoi(env);
  if (nGpus < 1) nGpus = 1;
  if (available >= 1 && nGpus > available) nGpus = available;
  if (lokib200_run_setup("Input", setupFile.c_str(), "Output", nGpus, 0, 1, nullptr) != 0) {
    const char* msg = lokib200_run_last_error();
    if (FILE* f = std::fopen("errorLog.txt", "a")) { std::fprintf(f, "Program stopped due to the following error:\n%s\n", msg); std::fclose(f); }
    std::printf("\033[31mProgram stopped due to the following error:\n%s\n\033[0m", msg);
    return EXIT_FAILURE;
  }
  return 0;
}
