// cloud_info.hpp — ROS1 wire image of rolo/CloudInfoStamp (msg/CloudInfoStamp.msg:1-28), host side.
//
// The three front-end stages of the reference hand a CloudInfoStamp message from node to node:
//   imageProjection  publishes index arrays + cloud_projected          (imageProjection.cpp:507-512)
//   featureExtraction clears the index arrays (freeCloudInfoMemory) and adds extracted_corner / extracted_surface /
//                     extracted_normal (always empty: the pushes are commented out)   (featureExtraction.cpp:268-287)
//   lidarOdometry    reads extracted_corner + extracted_surface (+ cloud_projected)   (lidarOdometry.cpp:521-524)
// Here the stages stay on the device (rg_front_end_frame); this file only produces / parses the byte layout at the ROS
// edge. Layout = roscpp serialisation (little endian; string / array = uint32 length + payload; time = 2 x uint32;
// bool = uint8) of the message, with every PointCloud2 as pcl::toROSMsg writes a pcl::PointCloud<pcl::PointXYZI>:
// fields x@0 y@4 z@8 intensity@16 (FLOAT32 = 7, count 1), point_step 32, height 1, width n, row_step 32 n, is_dense 1,
// is_bigendian 0. Point records are the 32-byte PointXYZI struct: x y z 1.0f | intensity 0 0 0 (the reference leaves the
// three floats after intensity uninitialised; zeros here).
#pragma once
#include <cstdint>
#include <cstring>
#include <string>

namespace rolo {
namespace wire {

constexpr uint32_t POINT_STEP = 32;
constexpr uint8_t FLOAT32 = 7;

struct Writer {
  uint8_t* p;        // null => size-only pass
  size_t cap, off;
  bool overflow;
  Writer(uint8_t* out, size_t capacity) : p(out), cap(capacity), off(0), overflow(false) {}
  void raw(const void* src, size_t n) {
    if (p) { if (off + n <= cap) std::memcpy(p + off, src, n); else overflow = true; }
    off += n;
  }
  size_t reserve(size_t n) { const size_t o = off; if (p && off + n > cap) overflow = true; off += n; return o; }
  void u8(uint8_t v) { raw(&v, 1); }
  void u32(uint32_t v) { raw(&v, 4); }
  void f32(float v) { raw(&v, 4); }
  void str(const char* s, size_t n) { u32((uint32_t)n); raw(s, n); }
};

struct Header { uint32_t seq, sec, nsec; const char* frame; size_t frame_len; };

inline void write_header(Writer& w, const Header& h) { w.u32(h.seq); w.u32(h.sec); w.u32(h.nsec); w.str(h.frame, h.frame_len); }

// a default-constructed sensor_msgs/PointCloud2 (what a never-assigned message member serialises to)
inline void write_empty_pointcloud2(Writer& w) {
  const Header z{0, 0, 0, "", 0};
  write_header(w, z);
  w.u32(0); w.u32(0);          // height, width
  w.u32(0);                    // fields[]
  w.u8(0); w.u32(0); w.u32(0); // is_bigendian, point_step, row_step
  w.u32(0);                    // data[]
  w.u8(0);                     // is_dense
}

// pcl::toROSMsg(PointCloud<PointXYZI>) with header.stamp / frame_id overwritten by publishCloud (utility.h:435-444);
// returns the offset of the n*32 data bytes (reserved, not written)
inline size_t write_xyzi_pointcloud2(Writer& w, const Header& h, uint32_t n) {
  const Header ph{0, h.sec, h.nsec, h.frame, h.frame_len};   // pcl header seq is 0
  write_header(w, ph);
  w.u32(1); w.u32(n);
  w.u32(4);
  const struct { const char* name; uint32_t off; } f[4] = {{"x", 0}, {"y", 4}, {"z", 8}, {"intensity", 16}};
  for (int i = 0; i < 4; i++) { w.str(f[i].name, std::strlen(f[i].name)); w.u32(f[i].off); w.u8(FLOAT32); w.u32(1); }
  w.u8(0); w.u32(POINT_STEP); w.u32(POINT_STEP * n);
  w.u32(POINT_STEP * n);
  const size_t data = w.reserve((size_t)POINT_STEP * n);
  w.u8(1);
  return data;
}

struct Scalars {
  float startOrientation, endOrientation, orientationDiff;
  float initialGuess[6];           // X Y Z Roll Pitch Yaw
  const float* covariance; uint32_t n_covariance;
  uint8_t odomAvailable;
};

struct Sections {                  // byte offsets of the device-produced payloads inside the message
  size_t start_ring, end_ring, col_ind, point_range;   // stage 0 only (SIZE_MAX otherwise)
  size_t projected, corner, surface;                   // point data (corner/surface: stage 1 only)
};

// stage 0: as published by imageProjection; stage 1: as published by featureExtraction
inline size_t write_cloud_info(Writer& w, int stage, const Header& h, const Scalars& s, uint32_t n_scan, uint32_t n_extracted,
                               uint32_t n_corner, uint32_t n_surface, Sections& sec) {
  write_header(w, h);
  sec.start_ring = sec.end_ring = sec.col_ind = sec.point_range = sec.corner = sec.surface = SIZE_MAX;
  if (stage == 0) {
    w.u32(n_scan); sec.start_ring = w.reserve(4ull * n_scan);
    w.u32(n_scan); sec.end_ring = w.reserve(4ull * n_scan);
    w.u32(n_extracted); sec.col_ind = w.reserve(4ull * n_extracted);
    w.u32(n_extracted); sec.point_range = w.reserve(4ull * n_extracted);
  } else {
    w.u32(0); w.u32(0); w.u32(0); w.u32(0);
  }
  w.f32(s.startOrientation); w.f32(s.endOrientation); w.f32(s.orientationDiff);
  for (int i = 0; i < 6; i++) w.f32(s.initialGuess[i]);
  w.u32(s.n_covariance);
  w.raw(s.covariance, 4ull * s.n_covariance);
  w.u8(s.odomAvailable);
  sec.projected = write_xyzi_pointcloud2(w, h, n_extracted);
  if (stage == 0) {
    write_empty_pointcloud2(w); write_empty_pointcloud2(w); write_empty_pointcloud2(w);
  } else {
    sec.corner = write_xyzi_pointcloud2(w, h, n_corner);
    sec.surface = write_xyzi_pointcloud2(w, h, n_surface);
    write_xyzi_pointcloud2(w, h, 0);            // extracted_normal: empty cloud through toROSMsg
  }
  write_empty_pointcloud2(w);                   // extracted_ground is never assigned
  return w.off;
}

// ---- parsing (consumer edge) ----
struct Reader {
  const uint8_t* p; size_t len, off; bool bad;
  Reader(const uint8_t* d, size_t n) : p(d), len(n), off(0), bad(false) {}
  bool need(size_t n) { if (bad || off + n > len || off + n < off) { bad = true; return false; } return true; }
  uint32_t u32() { uint32_t v = 0; if (need(4)) { std::memcpy(&v, p + off, 4); off += 4; } return v; }
  uint8_t u8() { uint8_t v = 0; if (need(1)) { v = p[off]; off += 1; } return v; }
  size_t skip(size_t n) { const size_t o = off; if (need(n)) off += n; return o; }
  std::string str() { const uint32_t n = u32(); std::string s; if (need(n)) { s.assign((const char*)p + off, n); off += n; } return s; }
};

struct CloudView {
  uint32_t n_points, point_step;
  int off_x, off_y, off_z, off_intensity;   // -1 when the field is missing
  size_t data;                              // offset of the point data in the message
};

inline void read_header(Reader& r) { r.u32(); r.u32(); r.u32(); r.str(); }

inline bool read_pointcloud2(Reader& r, CloudView& v) {
  read_header(r);
  const uint32_t height = r.u32(), width = r.u32();
  const uint32_t nf = r.u32();
  v.off_x = v.off_y = v.off_z = v.off_intensity = -1;
  for (uint32_t i = 0; i < nf && !r.bad; i++) {
    const std::string name = r.str();
    const uint32_t off = r.u32(); const uint8_t type = r.u8(); const uint32_t count = r.u32();
    if (type != FLOAT32 || count != 1) continue;
    if (name == "x") v.off_x = (int)off; else if (name == "y") v.off_y = (int)off; else if (name == "z") v.off_z = (int)off;
    else if (name == "intensity") v.off_intensity = (int)off;
  }
  const uint8_t big = r.u8();
  v.point_step = r.u32();
  r.u32();
  const uint32_t nbytes = r.u32();
  v.data = r.skip(nbytes);
  r.u8();
  v.n_points = height * width;
  if (r.bad || big) return false;
  if (v.n_points && ((uint64_t)v.n_points * v.point_step > nbytes)) return false;
  return true;
}

struct Parsed { uint32_t n_start, n_end, n_col, n_range; size_t start_ring, end_ring, col_ind, point_range; CloudView projected, corner, surface, normal, ground; };

inline bool parse_cloud_info(const uint8_t* msg, size_t len, Parsed& o) {
  Reader r(msg, len);
  read_header(r);
  o.n_start = r.u32(); o.start_ring = r.skip(4ull * o.n_start);
  o.n_end = r.u32(); o.end_ring = r.skip(4ull * o.n_end);
  o.n_col = r.u32(); o.col_ind = r.skip(4ull * o.n_col);
  o.n_range = r.u32(); o.point_range = r.skip(4ull * o.n_range);
  r.skip(4 * 9);
  const uint32_t ncov = r.u32(); r.skip(4ull * ncov);
  r.u8();
  return read_pointcloud2(r, o.projected) && read_pointcloud2(r, o.corner) && read_pointcloud2(r, o.surface) &&
         read_pointcloud2(r, o.normal) && read_pointcloud2(r, o.ground) && !r.bad && r.off == len;
}

}  // namespace wire
}  // namespace rolo
