/*
 * sREMD_cuda.c — replica-exchange launcher for the drop-in driver (the role of the reference's sREMD, REMDExec.c:28-160).
 *
 * Same command line and configuration file as sREMD:
 *     sREMD_cuda -f REMDConfig.txt -args <sDMD arguments without -REMD> -end [-seed N] [-gpus G] [-exe path/to/sDMD_cuda] [-server]
 * It forks one sDMD_cuda process per replica with the arguments sREMD would pass (-REMD localhost <port> <T> <num> <rate> <id>,
 * REMDExec.c:111-119; temperatures in Kelvin x 0.0019872041 printed with %lf) and places replica id on GPU id % G
 * (SDMD_CUDA_DEVICE).  What differs is the exchange:
 *
 *   default   no server process.  The replicas form one NCCL communicator (rank = replica id; rank 0 publishes the ncclUniqueId
 *             through a file the launcher names) and every exchange is sdmd_remd_exchange_energy: CalPotenE on the device ->
 *             ncclAllGather of {E, T, num} over NVLink -> the Server.c:145-192 sweep replicated on every rank with the server's
 *             glibc rand() stream (seed: -seed, default time(NULL) like REMDExec.c:141).  Needs one GPU per replica.
 *   -server   the reference's own sServer over TCP (started from the launcher's directory, like sREMD does), replicas may share
 *             GPUs; the start-up stagger sleep(1 + id) of REMDExec.c:133 is kept because the server sweeps in accept order.
 *
 * The child's stdout goes to outputData_<id>.txt like the reference's.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <fcntl.h>
#include <time.h>
#include <libgen.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include "sdmd_cuda.h"

#define CHAR_LENGTH 1024
#define BOLTZMANN 0.0019872041 /* kcal / mol / K, REMDExec.c:19 */
#define MAX_ARGS 48

static void usage(const char *me) {
    printf("Usage: %s -f [configuration file] -args [args of sDMD_cuda without flag -REMD] -end [-seed N] [-gpus G] [-exe path] [-server]\n", me);
    exit(EXIT_SUCCESS);
}

int main(int argc, char *argv[]) {
    char config[CHAR_LENGTH] = "", exe[CHAR_LENGTH] = "", dir[CHAR_LENGTH], buffer[CHAR_LENGTH], word[CHAR_LENGTH];
    char *command[MAX_ARGS];
    int nargs = 0, use_server = 0, gpus = 0;
    unsigned seed = (unsigned)time(NULL);

    for (int n = 1; n < argc; n++) {
        if (!strcmp(argv[n], "-h") || !strcmp(argv[n], "-help")) usage(argv[0]);
        else if (!strcmp(argv[n], "-f") && n + 1 < argc) snprintf(config, sizeof(config), "%s", argv[++n]);
        else if (!strcmp(argv[n], "-seed") && n + 1 < argc) seed = (unsigned)strtoul(argv[++n], NULL, 10);
        else if (!strcmp(argv[n], "-gpus") && n + 1 < argc) gpus = atoi(argv[++n]);
        else if (!strcmp(argv[n], "-exe") && n + 1 < argc) snprintf(exe, sizeof(exe), "%s", argv[++n]);
        else if (!strcmp(argv[n], "-server")) use_server = 1;
        else if (!strcmp(argv[n], "-args")) {
            n++;
            while (n < argc && strcmp(argv[n], "-end")) {
                if (nargs >= MAX_ARGS - 9) { printf("!!ERROR!!: too many arguments after -args\n"); exit(EXIT_FAILURE); }
                command[nargs++] = argv[n++];
            }
        }
    }
    if (!config[0] || nargs == 0) usage(argv[0]);
    /* sREMD looks for its executables in the working directory (REMDExec.c:77-78); here: next to the launcher */
    snprintf(buffer, sizeof(buffer), "%s", argv[0]);
    snprintf(dir, sizeof(dir), "%s", dirname(buffer));
    if (!exe[0]) snprintf(exe, sizeof(exe), "%s/sDMD_cuda", dir);
    if (gpus <= 0) gpus = sdmd_cuda_device_count();
    if (gpus <= 0) { printf("!!ERROR!!: no CUDA device\n"); exit(EXIT_FAILURE); }

    /* REMDConfig.txt, read exactly like REMDExec.c:80-109 */
    FILE *cf = fopen(config, "r");
    if (!cf) { printf("!!ERROR!!: cannot find the configuration file at %s\n", config); exit(EXIT_FAILURE); }
    int replicaNum = 0;
    char portNo[32], exRate[32];
    fgets(buffer, CHAR_LENGTH, cf);
    fscanf(cf, "%s%s%i\n", word, word, &replicaNum);
    double *T = (double *)calloc(replicaNum > 0 ? replicaNum : 1, sizeof(double));
    fgets(buffer, CHAR_LENGTH, cf);
    int pos = 0, cur = 0;
    sscanf(buffer, "%s%s%n", word, word, &cur);
    pos += cur;
    for (int i = 0; i < replicaNum; i++) {
        if (sscanf(buffer + pos, "%lf%n", &T[i], &cur) != 1) {
            printf("!!ERROR!!: the number of temperatures provided is less than the number of replica!\n");
            exit(EXIT_FAILURE);
        }
        T[i] *= BOLTZMANN;
        pos += cur;
    }
    double extra;
    if (sscanf(buffer + pos, "%lf", &extra) == 1) {
        printf("!!ERROR!!: the number of temperatures provided is larger than the number of replica!\n");
        exit(EXIT_FAILURE);
    }
    fscanf(cf, "%s%s%31s", word, word, portNo);
    fscanf(cf, "%s%s%31s", word, word, exRate);
    fclose(cf);
    if (!use_server && replicaNum > gpus) {
        printf("!!ERROR!!: %i replicas on %i GPU(s): the NCCL exchange needs one GPU per replica (use -server to share GPUs)\n", replicaNum, gpus);
        exit(EXIT_FAILURE);
    }

    char idfile[CHAR_LENGTH];
    snprintf(idfile, sizeof(idfile), "/tmp/sdmd_remd_nccl_id_%ld_%u", (long)getpid(), seed);
    unlink(idfile);
    char sT[64], sNum[16], sId[16], sRanks[16], sSeed[32], sDev[16];
    snprintf(sRanks, sizeof(sRanks), "%i", replicaNum);
    snprintf(sSeed, sizeof(sSeed), "%u", seed);
    pid_t server_pid = -1;
    for (int id = 0; id < replicaNum; id++) {
        pid_t child = fork();
        if (child == 0) {
            char outputData[64];
            snprintf(outputData, sizeof(outputData), "outputData_%i.txt", id);
            int fd = open(outputData, O_RDWR | O_CREAT | O_TRUNC, S_IRUSR | S_IWUSR);
            dup2(fd, 1);
            close(fd);
            snprintf(sT, sizeof(sT), "%lf", T[id]);          /* REMDExec.c:118 */
            snprintf(sNum, sizeof(sNum), "%i", id + 1);
            snprintf(sId, sizeof(sId), "%i", id);
            snprintf(sDev, sizeof(sDev), "%i", id % gpus);
            char *cmd[MAX_ARGS];
            int k = 0;
            cmd[k++] = exe;
            for (int a = 0; a < nargs; a++) cmd[k++] = command[a];
            cmd[k++] = "-REMD"; cmd[k++] = "localhost"; cmd[k++] = portNo; cmd[k++] = sT; cmd[k++] = sNum; cmd[k++] = exRate; cmd[k++] = sId;
            cmd[k] = NULL;
            setenv("SDMD_CUDA_DEVICE", sDev, 1);
            if (!use_server) {
                setenv("SDMD_REMD_NCCL", "1", 1);
                setenv("SDMD_REMD_NRANKS", sRanks, 1);
                setenv("SDMD_REMD_RANK", sId, 1);
                setenv("SDMD_REMD_ID_FILE", idfile, 1);
                setenv("SDMD_REMD_SEED", sSeed, 1);
            } else {
                sleep(1 + id);                               /* REMDExec.c:133: connect in replica order */
            }
            if (execv(exe, cmd) < 0) perror("!!ERROR!!");
            exit(EXIT_FAILURE);
        }
    }
    if (use_server) {
        server_pid = fork();
        if (server_pid == 0) {
            char execServer[CHAR_LENGTH];
            snprintf(execServer, sizeof(execServer), "%s/sServer", dir);
            char *serverCommand[] = {"./sServer", portNo, sRanks, sSeed, NULL};
            if (execv(execServer, serverCommand) < 0) perror("!!ERROR!!");
            exit(EXIT_FAILURE);
        }
    }
    int status = 0, failed = 0;
    pid_t w;
    while ((w = wait(&status)) > 0)
        if (w != server_pid && (!WIFEXITED(status) || WEXITSTATUS(status) != 0)) failed++;
    unlink(idfile);
    free(T);
    if (failed) { printf("!!ERROR!!: %i replica process(es) failed\n", failed); return EXIT_FAILURE; }
    return 0;
}
