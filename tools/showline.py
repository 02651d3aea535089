import json, sys
tag = sys.argv[1]
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
s = d["stats_last_step"]
print("%s value=%.3e e2e=%.3e kernel_ms=%.3f step_ms=%.3f e2e_ms=%.3f steps=%s | seed=%.2f search=%.2f rows_sort=%.2f nb=%.2f fin=%.2f build=%.2f rounds=%d rays=%d dup=%d c32=%d c64=%d retries=%d" % (
    tag, d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms_per_step"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["step_ms_list"],
    s["ms_seed"], s["ms_search"], s["ms_rows_sort"], s["ms_neighbors"], s["ms_finalize"], s["ms_build"], s["rounds"], s["raycasts"],
    s["duplicate_hits"], s["candidates_fp32"], s["candidates_fp64"], s["capacity_retries"]))
