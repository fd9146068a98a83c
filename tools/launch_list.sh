#!/bin/bash
# usage: tools/launch_list.sh <out.csv> <command...>  -- per-launch durations of this library's kernels under ncu
out=$1; shift
RX='pack_reads|map_reads|tally_units|dict_|set_i64|add_i64|em_loop|em_class|em_tx|em_decide|multinomial|counts_to_plan|broadcast_kernel|tpm_finish|transpose_kernel|sum_counts|expand_rows|iota_|histogram_kernel|gather_i32|fill_i32|eff_len|map_kmers|select_heavy|class_sort_key|permuted_|ordered_|compact_cols|extract_cols|newlines|fastq_units'
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$RX" -c 4000 --csv --log-file "$out" "$@"
