// fastkmer_cli — command-line twin of skc.test.LocalTestKmerCounter / TestKmerCounter
// (LTKC:18-56, TKC:15-54): same positional arguments, in the CODE's order
//     k m x B useHT sequenceType inputPath outputPath prefix write enableKryo useCustomPartitioner [numPartitionTasks]
// (LTKC:35-48; the README's parameter table lists useHT before B, the code does not),
// same derived b and output directory (TCFG:32-33), same bin files.
#include "../../include/fastkmer_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

int main(int argc, char** argv) {
    if (argc < 13) {
        fprintf(stderr, "usage: %s k m x B useHT sequenceType inputPath outputPath prefix write enableKryo useCustomPartitioner [numPartitionTasks]\n", argv[0]);
        return 2;
    }
    fkm_config c;
    memset(&c, 0, sizeof c);
    c.k = atoi(argv[1]); c.m = atoi(argv[2]); c.x = atoi(argv[3]); c.max_b = atoi(argv[4]);
    c.use_ht = atoi(argv[5]) == 1; c.sequence_type = atoi(argv[6]);
    c.dataset = argv[7]; c.output_directory = argv[8]; c.prefix = argv[9];
    c.write = atoi(argv[10]) == 1; c.use_kryo_serializer = atoi(argv[11]) == 1;
    c.use_custom_partitioner = atoi(argv[12]) == 1;
    if (c.use_custom_partitioner) {
        if (argc < 14) { fprintf(stderr, "numPartitionTasks missing\n"); return 2; }
        c.num_partition_tasks = atoi(argv[13]);
    }
    // FKM_GPUS=N: the job runs on GPUs 0..N-1 of this node (the reference spreads it over its Spark executors)
    const int n_gpus = getenv("FKM_GPUS") ? atoi(getenv("FKM_GPUS")) : 1;
    fkm_ctx* ctx = nullptr;
    fkm_stats st;
    int rc;
    if (n_gpus > 1) {
        int32_t dev[64];
        for (int i = 0; i < n_gpus && i < 64; i++) dev[i] = i;
        rc = fkm_execute_job_multi(dev, n_gpus < 64 ? n_gpus : 64, &c, &st);
    } else {
        if (fkm_ctx_create(-1, nullptr, &ctx) != FKM_OK) { fprintf(stderr, "fastkmer_cli: %s\n", fkm_last_error()); return 1; }
        rc = fkm_execute_job(ctx, &c, &st);
    }
    if (rc != FKM_OK) { fprintf(stderr, "fastkmer_cli: %s\n", fkm_last_error()); fkm_ctx_destroy(ctx); return 1; }
    char dir[4096]; int32_t b = 0;
    fkm_derive(&c, &b, dir, sizeof dir);
    printf("Kmer counting on B200.\nk: %d\nm: %d\nx: %d\nb: %d\nSequence type: %d\nUsing HT: %d\nWriting: %d\n", c.k, c.m, c.x, b, c.sequence_type, c.use_ht, c.write);
    printf("bases %llu kmers %llu distinct %llu superkmers %llu bins %llu | device %.3f ms total %.3f ms | out %s\n",
           (unsigned long long)st.n_bases, (unsigned long long)st.n_kmers, (unsigned long long)st.n_distinct,
           (unsigned long long)st.n_superkmers, (unsigned long long)st.n_nonempty_bins, st.ms_stage[7], st.ms_total, c.write ? dir : "(write=0)");
    fkm_ctx_destroy(ctx);
    return 0;
}
