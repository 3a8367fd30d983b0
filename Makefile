# Top-level conveniences. `make` builds everything in-tree for sm_100a (same as __graft_entry__.build()).
PY ?= python
all:
	$(PY) -c "import __graft_entry__ as g; g.build()"

test-cpu:
	$(PY) -m pytest tests/ -x -q -m "not gpu"

test-gpu:
	$(PY) -m pytest tests/ -x -q -m gpu

# compute-sanitizer over the corner-case / flag-forced device paths (needs a GPU; on the B200 box: gpurun -- make sanitize).
# GRLGPU_NO_POOL=1 gives every buffer its own cudaMalloc so memcheck sees true bounds; racecheck covers the shared-memory
# publish protocol of dedup_cached_kernel and the radix / scan staging. Logs land in gpurun_out/ (copy into profiles/).
SAN_OUT ?= gpurun_out
sanitize:
	mkdir -p $(SAN_OUT)
	GRLGPU_NO_POOL=1 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $(SAN_OUT)/sanitize_memcheck.log $(PY) tests/sanitize_driver.py $(SAN_ARGS)
	compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 --log-file $(SAN_OUT)/sanitize_racecheck.log $(PY) tests/sanitize_driver.py $(SAN_ARGS)
	compute-sanitizer --tool synccheck --error-exitcode 9 --log-file $(SAN_OUT)/sanitize_synccheck.log $(PY) tests/sanitize_driver.py $(SAN_ARGS)
	tail -n 3 $(SAN_OUT)/sanitize_memcheck.log $(SAN_OUT)/sanitize_racecheck.log $(SAN_OUT)/sanitize_synccheck.log

# memcheck over the multi-rank rounds and the packed output only (seconds; profiles/r02_sanitize_memcheck_quick_final.log)
sanitize-quick:
	mkdir -p $(SAN_OUT)
	GRLGPU_NO_POOL=1 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $(SAN_OUT)/sanitize_memcheck_quick.log $(PY) tests/sanitize_driver.py --quick

.PHONY: all test-cpu test-gpu sanitize sanitize-quick
