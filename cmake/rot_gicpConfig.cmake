# rot_gicpConfig.cmake — CMake package of the B200-native rot_gicp replacement.
#
#   list(APPEND CMAKE_PREFIX_PATH <this repository>/cmake)      # or -Drot_gicp_DIR=<this repository>/cmake
#   find_package(rot_gicp REQUIRED)
#   target_link_libraries(${PROJECT_NAME}_lidarOdometry ... rot_gicp)     # the reference's own link line (CMakeLists.txt:189) stays as it is
#
# Defines the imported target `rot_gicp` (the name the reference's add_library(rot_gicp SHARED ...) creates, CMakeLists.txt:160-176): the
# header tree with the reference's include paths (<rot_gicp/gicp/rot_vgicp.hpp>, <rot_gicp/gicp/lsq_registration.hpp>, ...) and
# librot_gicp.so, which is librolo_gicp_b200.so under the reference's library name (built by `python -m rolo_b200.build`).
# In the reference's CMakeLists.txt remove the add_library(rot_gicp ...) block (lines 160-176) and add the two lines above.
get_filename_component(_rot_gicp_root "${CMAKE_CURRENT_LIST_DIR}/.." ABSOLUTE)
set(rot_gicp_INCLUDE_DIRS "${_rot_gicp_root}/include")
set(rot_gicp_LIBRARY "${_rot_gicp_root}/rolo_b200/lib/librot_gicp.so")
if(NOT EXISTS "${rot_gicp_LIBRARY}")
  message(FATAL_ERROR "rot_gicp: ${rot_gicp_LIBRARY} not found — build it first: python -m rolo_b200.build")
endif()
if(NOT TARGET rot_gicp)
  add_library(rot_gicp SHARED IMPORTED GLOBAL)
  set_target_properties(rot_gicp PROPERTIES
    IMPORTED_LOCATION "${rot_gicp_LIBRARY}"
    INTERFACE_INCLUDE_DIRECTORIES "${rot_gicp_INCLUDE_DIRS}")
endif()
set(rot_gicp_LIBRARIES rot_gicp)
set(rot_gicp_FOUND TRUE)
