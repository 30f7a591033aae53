#!/bin/bash
O=gpurun_out
python tools/timeline.py --isolated > $O/r2o_timeline.jsonl 2>> $O/r2o_err.log
python tools/timeline.py --isolated --envs 16384 >> $O/r2o_timeline.jsonl 2>> $O/r2o_err.log
python tools/timeline.py >> $O/r2o_timeline.jsonl 2>> $O/r2o_err.log
python tools/timeline.py --isolated --envs 2048 >> $O/r2o_timeline.jsonl 2>> $O/r2o_err.log
