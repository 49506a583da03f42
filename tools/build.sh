#!/bin/bash
# build libspwombling.so in-tree, whatever the current directory is
cd "$(dirname "$0")/.." && python -m spwombling_b200.build "$@"
