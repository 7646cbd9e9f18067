# Snakemake fragment: Phylign's match stage on B200.
#
# Include it from the reference Snakefile after its `batches`, `cobs_dir`, `decompression_dir` and
# `config` definitions:
#
#   include: "/path/to/phylign-b200/integration/match_gpu.smk"
#
# and select it in config.yaml with the reference's own switch (config.yaml:104, Snakefile:124-130):
#
#   index_load_mode: gpu          # beside the reference's mem-stream | mem-disk | mmap-disk
#   phylign_b200_dir: /path/to/phylign-b200
#   gpus: 1                       # GPUs of the node the rule runs on (--gpus N: one worker per GPU)
#
# With `index_load_mode: gpu` the rule below takes precedence over decompress_cobs / run_cobs /
# decompress_and_run_cobs (Snakefile:364-487) and translate_matches (Snakefile:490-520); with any
# other value this file changes nothing.  get_index_load_mode() of the reference asserts the value
# against its own list (Snakefile:124-130): add "gpu" to `allowed_index_load_modes` there (one word;
# INTEGRATION.md section 3 shows the diff).
#
# File contracts kept: intermediate/03_match/{batch}____{qfile}.gz, intermediate/04_filter/{qfile}.fa,
# logs/benchmarks/run_cobs/{batch}____{qfile}.txt (scripts/benchmark.py format family).
# keep_cobs_indexes / decompression_dir keep their meaning: decompressed {batch}.cobs_classic files
# found in {decompression_dir} are loaded instead of the .xz (no LZMA decode, PCIe-speed load), and
# with keep_cobs_indexes: True they are left there while the .xz streams in.

PHYLIGN_B200 = config.get("phylign_b200_dir", "/path/to/phylign-b200")
GPU_MODE = config.get("index_load_mode", "mem-stream") == "gpu"

if GPU_MODE:

    ruleorder: match_gpu > decompress_and_run_cobs > run_cobs > translate_matches

    rule match_gpu:
        """COBS matching of all batches + top-N merge in one resident GPU job"""
        output:
            fa="intermediate/04_filter/{qfile}.fa",
            matches=[f"intermediate/03_match/{batch}____{{qfile}}.gz" for batch in batches],
        input:
            fa="intermediate/01_queries_merged/{qfile}.fa",
            xz=[f"{cobs_dir}/{batch}.cobs_classic.xz" for batch in batches],
            decompressed_indexes_sizes="data/decompressed_indexes_sizes.txt",
        threads: workflow.cores          # xz decoders, file readers and gzip writers use the host cores
        resources:
            gpu=int(config.get("gpus", 1)),
        params:
            kmer_thres=config["cobs_kmer_thres"],
            nb_best_hits=config["nb_best_hits"],
            batches_fn=config["batches"],
            gpus=int(config.get("gpus", 1)),
            decompression_dir=decompression_dir,
            keep="--keep-cobs-indexes" if config.get("keep_cobs_indexes", False) else "",
        log:
            "logs/03_match_gpu/{qfile}.log",
        shell:
            """
            ./scripts/benchmark.py --log logs/benchmarks/match_gpu/match_gpu___{wildcards.qfile}.txt \\
                'PYTHONPATH={PHYLIGN_B200} python3 -m phylign_b200.cli match-db \\
                        --cobs-dir {cobs_dir} --batches {params.batches_fn} \\
                        -q {input.fa} --qfile {wildcards.qfile} \\
                        --match-dir intermediate/03_match --filter-out {output.fa} \\
                        -t {params.kmer_thres} -n {params.nb_best_hits} \\
                        --index-sizes-table {input.decompressed_indexes_sizes} \\
                        --decompression-dir {params.decompression_dir} {params.keep} \\
                        --benchmark-dir logs/benchmarks/run_cobs \\
                        --gpus {params.gpus} --load-workers {threads} --resume 2>{log}'
            """
