# ncu --set full summaries of the dominant kernel of each 8(f) row: bash tools/final_rows_ncu.sh <tag>   (under gpurun)
tag=${1:-r2z}
mkdir -p gpurun_out
capture() {   # capture <name> <kernel regex> <skip> <row>
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c 1 -o gpurun_out/${tag}_$1 python tools/run_rows.py $4 > gpurun_out/${tag}_ncu_$1.log 2>&1
    ncu -i gpurun_out/${tag}_$1.ncu-rep --page raw --csv > gpurun_out/${tag}_$1.raw.csv 2>/dev/null
    python tools/summarize_ncu.py raw gpurun_out/${tag}_$1.raw.csv gpurun_out/${tag}_${1}_ncu_full.csv "$2" > /dev/null 2>&1
    rm -f gpurun_out/${tag}_$1.ncu-rep gpurun_out/${tag}_$1.raw.csv
    grep -E "^(gpu__time_duration.sum|dram__bytes_read.sum|dram__bytes_write.sum|sm__issue_active.avg.pct_of_peak_sustained_elapsed|smsp__thread_inst_executed_per_inst_executed.ratio|launch__registers_per_thread|lts__t_sector_hit_rate.pct|l1tex__t_sector_hit_rate.pct|dram__throughput.avg.pct_of_peak_sustained_elapsed)," gpurun_out/${tag}_${1}_ncu_full.csv | tr '\n' ' '; echo " <- $1"
}
capture normals_ball_kernel normals_ball_kernel 1 ball
capture sort_scatter_kernel sort_scatter_kernel 12 dedup
capture voxel_insert_kernel voxel_insert_kernel 1 voxel
capture sinkhorn_half_kernel sinkhorn_half_kernel 2 sinkhorn
