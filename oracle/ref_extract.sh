#!/bin/sh
# Cuts the hot-path function bodies out of reference files that also include ROS headers, by line range,
# into _ref/extract/ (removed again after the compile; see Makefile target _ref_hotpath).  Each range is
# checked against the text its first line must contain, so a changed reference fails loudly.
# usage: ref_extract.sh <reference root> <output dir>
set -e
REF="$1"; OUT="$2"
mkdir -p "$OUT"
cut() {  # file first last out expected-first-line-fragment
  sed -n "$2,$3p" "$REF/$1" > "$OUT/$4"
  if ! head -1 "$OUT/$4" | grep -q -F -- "$5"; then
    echo "reference changed: $1:$2 does not contain '$5'" >&2; exit 1
  fi
}
cut nbveplanner/src/rrt.cpp 163 264 rrt_cpp_163_264.inc 'void RrtTree::iterate() {'
cut nbveplanner/src/rrt.cpp 351 479 rrt_cpp_351_479.inc 'double RrtTree::optimized_gain(Pose &state) {'
cut nbveplanner/src/rrt.cpp 481 579 rrt_cpp_481_579.inc 'double RrtTree::gain(Pose &state) {'
cut nbveplanner/src/voxblox_manager.cpp 14 33 voxblox_manager_cpp_14_33.inc 'VoxelStatus VoxbloxManager::getVoxelStatusByIndex('
cut nbveplanner/src/voxblox_manager.cpp 121 129 voxblox_manager_cpp_121_129.inc 'bool HighResManager::checkCollisionWithRobotAtVoxel('
cut nbveplanner/include/nbveplanner/voxblox_manager.h 35 37 voxblox_manager_h_35_37.inc 'double getResolution() const'
cut nbveplanner/include/nbveplanner/voxblox_manager.h 89 108 voxblox_manager_h_89_108.inc 'template <typename T>'
cut nbveplanner/src/voxblox_manager.cpp 73 90 voxblox_manager_cpp_73_90.inc 'VoxelStatus LowResManager::getVoxelStatus(const Point &position) const {'
cut nbveplanner/src/history.cpp 96 137 history_cpp_96_137.inc 'std::string convertToString(const Eigen::Vector3d &vec) {'
cut voxblox/voxblox/src/core/esdf_map.cc 22 52 esdf_map_cc_22_52.inc 'bool EsdfMap::getDistanceAndGradientAtPosition('
