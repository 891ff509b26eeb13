OUT=gpurun_out/r03c; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/knn_launches.csv python tools/knn_profile_run.py > $OUT/launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_search -s 2 -c 1 -o $OUT/knn_search -f python tools/knn_profile_run.py > $OUT/full.log 2>&1
tail -3 $OUT/full.log; ls -la $OUT
