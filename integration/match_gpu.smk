# Snakemake fragment: Phylign's match stage on B200 (include from the reference Snakefile after its
# `batches`, `cobs_dir`, `config` definitions; replaces rules decompress_cobs / run_cobs /
# decompress_and_run_cobs (Snakefile:364-487) and translate_matches (Snakefile:490-520)).
#
#   include: "/path/to/phylign-b200/integration/match_gpu.smk"
#   ruleorder: match_gpu > decompress_and_run_cobs > run_cobs > translate_matches
#
# File contracts kept: intermediate/03_match/{batch}____{qfile}.gz, intermediate/04_filter/{qfile}.fa.

PHYLIGN_B200 = config.get("phylign_b200_dir", "/path/to/phylign-b200")


rule match_gpu:
    """COBS matching of all batches + top-N merge in one resident GPU job"""
    output:
        fa="intermediate/04_filter/{qfile}.fa",
        matches=[f"intermediate/03_match/{batch}____{{qfile}}.gz" for batch in batches],
    input:
        fa="intermediate/01_queries_merged/{qfile}.fa",
        xz=[f"{cobs_dir}/{batch}.cobs_classic.xz" for batch in batches],
        decompressed_indexes_sizes="data/decompressed_indexes_sizes.txt",
    threads: workflow.cores          # the xz decoders use the host cores; the GPU does the matching
    resources:
        gpu=1,
    params:
        kmer_thres=config["cobs_kmer_thres"],
        nb_best_hits=config["nb_best_hits"],
        batches_fn=config["batches"],
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
                    --load-workers {threads} --resume 2>{log}'
        """
