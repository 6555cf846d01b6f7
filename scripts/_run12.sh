cd $GRAFT_REPO_ROOT
timeout 300 python scripts/prof_scatter.py > gpurun_out/r2_prof_scatter.txt 2>&1; echo "scatter rc=$?"
timeout 300 python scripts/prof_spmm.py > gpurun_out/r2_prof_spmm.txt 2>&1; echo "spmm rc=$?"
timeout 300 python scripts/prof_resident.py gpurun_out/r2_prof_resident_v3.json > /dev/null 2>&1; echo "resident rc=$?"
tail -12 gpurun_out/r2_prof_scatter.txt; tail -12 gpurun_out/r2_prof_spmm.txt
