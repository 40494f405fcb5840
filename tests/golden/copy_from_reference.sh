#!/bin/sh
# Re-creates tests/golden/ from the reference tree (run in the build container, where /root/reference exists).
set -e
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
for i in 1 2 3 4 5 6; do cp -r "$REF/test/benchmark/data/benchmark_$i" "$HERE/"; done
mkdir -p "$HERE/registration_test"
for f in objectIn.csv sceneIn.csv rotation_only_src.csv translation_test_v1_inliers.csv translation_test_v2_inliers.csv; do
  cp "$REF/test/teaser/data/registration_test/$f" "$HERE/registration_test/"
done
cp "$REF/examples/example_data/bun_zipper_res3.ply" "$HERE/"
cp "$REF/test/teaser/data/cube.ply" "$HERE/"
cp "$REF/test/teaser/data/bunny.pcd" "$REF/test/teaser/data/bunny_fpfh.csv" "$HERE/"
cp "$REF/test/teaser/data/canstick.ply" "$REF/test/teaser/data/matcher-test-object-1.ply" "$REF/test/teaser/data/matcher-test-scene-1.ply" "$REF/test/teaser/data/matcher-test-matches-1.csv" "$HERE/"
cp -r "$REF/test/teaser/data/certification_small_instances" "$REF/test/teaser/data/certification_large_instances" "$HERE/"
chmod -R u+w "$HERE"
