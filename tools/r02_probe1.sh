#!/bin/bash
# round-2 GPU call 1: tests, producer-wait variants, CTA form for SimpleThreshold, ring geometry for under-filled launches, gather probe
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe1.txt
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== gather probe"; timeout 300 build/bin/gather_probe 7200 122880 1
echo "== producer wait variants"
for a in "5920 64 FIR 5" "1480 340 SimpleThreshold 60 wib2" "1480 340 FIR 5 wib2" "1480 340 AbsRS 60 wib2"; do
  tools/variant_probe.sh "$a" base pw0 pw1 pw2s50 pw2s500
done
echo "== CTA form, SimpleThreshold"
tools/variant_probe.sh "5920 64 SimpleThreshold 60" base qs5 qs6
echo "== ring geometry of the warp form"
for a in "40 2048 SimpleThreshold 60" "240 64 SimpleThreshold 60" "750 64 SimpleThreshold 60" "1500 64 SimpleThreshold 60" "3000 64 SimpleThreshold 60" "5920 64 SimpleThreshold 60"; do
  tools/variant_probe.sh "$a" base g3x32 g4x32 g8x32 g4x16 g8x16
done
} > $OUT 2>&1
tail -5 $OUT
