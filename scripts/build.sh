#!/bin/bash
# rebuild libequss_b200.so in-tree (same as __graft_entry__.build())
cd "$(dirname "$0")/.." && python expand-and-quantize-for-unsupervised-semantic-segmentation_b200/build.py "$@"
