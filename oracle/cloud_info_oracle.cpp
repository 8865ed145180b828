/*
 * cloud_info_oracle.cpp — CPU ORACLE (test infrastructure, NOT product code) for the rolo/CloudInfoStamp wire image.
 *
 * Restates, byte for byte, what the reference's nodes put on the wire when they publish the message that links the three
 * front-end stages:
 *   msg/CloudInfoStamp.msg:1-28                 field order
 *   src/imageProjection.cpp:507-512             stage 0: index arrays + cloud_projected = publishCloud(extractedCloud)
 *   src/featureExtraction.cpp:268-287           stage 1: freeCloudInfoMemory() + extracted_corner/surface/normal
 *   include/rolo/utility.h:435-444              publishCloud: pcl::toROSMsg, then stamp + frame_id overwritten
 *   src/lidarOdometry.cpp:521-524               consumer: fromROSMsg(corner), fromROSMsg(surface), featureLast = corner + surf
 * PARITY UNPINNED: roscpp's serialiser and pcl::toROSMsg / fromROSMsg are third-party code absent from /root/reference
 * (ROS Noetic roscpp_serialization, PCL >= 1.10 pcl_conversions; README.md:28-31, unpinned). Their published rules are
 * restated here: little endian; string and T[] = uint32 count + payload; time = uint32 sec, uint32 nsec; bool = uint8;
 * toROSMsg(PointCloud<PointXYZI>) = fields x@0,y@4,z@8,intensity@16 (datatype 7 = FLOAT32, count 1), point_step = 32 =
 * sizeof(PointXYZI), height 1, width n, row_step = 32 n, is_bigendian 0, is_dense 1, pcl header seq 0, data = the raw structs.
 * The struct's padding floats are written as {1.0f after z, 0 after intensity}; the reference leaves the latter
 * uninitialised, so those 12 bytes per point cannot be pinned by anything.
 */
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

namespace {
typedef std::vector<uint8_t> Buf;
template <class T> void put(Buf& b, T v) { const uint8_t* p = reinterpret_cast<const uint8_t*>(&v); b.insert(b.end(), p, p + sizeof(T)); }
void put_string(Buf& b, const std::string& s) { put<uint32_t>(b, (uint32_t)s.size()); b.insert(b.end(), s.begin(), s.end()); }
void put_std_header(Buf& b, uint32_t seq, uint32_t sec, uint32_t nsec, const std::string& frame) {
  put(b, seq); put(b, sec); put(b, nsec); put_string(b, frame);
}
template <class T> void put_array(Buf& b, const T* v, int n) { put<uint32_t>(b, (uint32_t)n); for (int i = 0; i < n; i++) put(b, v[i]); }

// a PointCloud2 member nobody assigned
void put_blank_cloud(Buf& b) {
  put_std_header(b, 0, 0, 0, "");
  put<uint32_t>(b, 0); put<uint32_t>(b, 0); put<uint32_t>(b, 0); put<uint8_t>(b, 0); put<uint32_t>(b, 0); put<uint32_t>(b, 0); put<uint32_t>(b, 0);
  put<uint8_t>(b, 0);
}
// publishCloud(cloud of n PointXYZI)
void put_xyzi_cloud(Buf& b, uint32_t sec, uint32_t nsec, const std::string& frame, const float* xyzi, int n) {
  put_std_header(b, 0, sec, nsec, frame);
  put<uint32_t>(b, 1); put<uint32_t>(b, (uint32_t)n);
  put<uint32_t>(b, 4);
  const char* names[4] = {"x", "y", "z", "intensity"};
  const uint32_t offs[4] = {0, 4, 8, 16};
  for (int f = 0; f < 4; f++) { put_string(b, names[f]); put(b, offs[f]); put<uint8_t>(b, 7); put<uint32_t>(b, 1); }
  put<uint8_t>(b, 0); put<uint32_t>(b, 32); put<uint32_t>(b, 32u * (uint32_t)n);
  put<uint32_t>(b, 32u * (uint32_t)n);
  for (int i = 0; i < n; i++) {
    const float rec[8] = {xyzi[4 * i], xyzi[4 * i + 1], xyzi[4 * i + 2], 1.0f, xyzi[4 * i + 3], 0.f, 0.f, 0.f};
    for (int k = 0; k < 8; k++) put(b, rec[k]);
  }
  put<uint8_t>(b, 1);
}

struct Cur {
  const uint8_t* p; size_t n, at; bool ok;
  template <class T> T get() { T v{}; if (at + sizeof(T) > n) { ok = false; return v; } std::memcpy(&v, p + at, sizeof(T)); at += sizeof(T); return v; }
  std::string get_string() { uint32_t k = get<uint32_t>(); if (at + k > n) { ok = false; return ""; } std::string s((const char*)p + at, k); at += k; return s; }
  void skip(size_t k) { if (at + k > n) ok = false; else at += k; }
};
// fromROSMsg into PointXYZI: x,y,z,intensity by field name
bool get_cloud(Cur& c, std::vector<float>& xyzi) {
  c.get<uint32_t>(); c.get<uint32_t>(); c.get<uint32_t>(); c.get_string();
  const uint32_t h = c.get<uint32_t>(), w = c.get<uint32_t>(), nf = c.get<uint32_t>();
  int off[4] = {-1, -1, -1, -1};
  for (uint32_t f = 0; f < nf && c.ok; f++) {
    const std::string name = c.get_string(); const uint32_t o = c.get<uint32_t>(); const uint8_t t = c.get<uint8_t>(); c.get<uint32_t>();
    const char* want[4] = {"x", "y", "z", "intensity"};
    for (int k = 0; k < 4; k++) if (name == want[k] && t == 7) off[k] = (int)o;
  }
  c.get<uint8_t>();
  const uint32_t step = c.get<uint32_t>(); c.get<uint32_t>();
  const uint32_t nbytes = c.get<uint32_t>();
  const size_t data = c.at;
  c.skip(nbytes); c.get<uint8_t>();
  if (!c.ok) return false;
  const size_t npts = (size_t)h * w;
  xyzi.assign(npts * 4, 0.f);
  for (size_t i = 0; i < npts; i++)
    for (int k = 0; k < 4; k++)
      if (off[k] >= 0) std::memcpy(&xyzi[i * 4 + k], c.p + data + i * step + off[k], 4);
  return true;
}
}  // namespace

extern "C" {

/* scalars9 = startOrientation, endOrientation, orientationDiff, initialGuess X Y Z Roll Pitch Yaw. Returns the message size;
 * writes it when out != NULL and cap suffices. */
size_t ro_cloud_info_serialize(int stage, uint32_t seq, uint32_t sec, uint32_t nsec, const char* frame,
                               const int32_t* start_ring, const int32_t* end_ring, int n_scan,
                               const int32_t* col_ind, const float* point_range, const float* extracted_xyzi, int n_extracted,
                               const float* corner_xyzi, int n_corner, const float* surface_xyzi, int n_surface,
                               const float* scalars9, const float* covariance, int n_cov, int odom_available,
                               uint8_t* out, size_t cap) {
  Buf b;
  const std::string fr(frame ? frame : "");
  put_std_header(b, seq, sec, nsec, fr);
  if (stage == 0) {
    put_array(b, start_ring, n_scan); put_array(b, end_ring, n_scan); put_array(b, col_ind, n_extracted); put_array(b, point_range, n_extracted);
  } else {
    for (int k = 0; k < 4; k++) put<uint32_t>(b, 0);
  }
  for (int k = 0; k < 9; k++) put(b, scalars9 ? scalars9[k] : 0.f);
  put_array(b, covariance, n_cov);
  put<uint8_t>(b, odom_available ? 1 : 0);
  put_xyzi_cloud(b, sec, nsec, fr, extracted_xyzi, n_extracted);                 // cloud_projected
  if (stage == 0) { put_blank_cloud(b); put_blank_cloud(b); put_blank_cloud(b); }
  else {
    put_xyzi_cloud(b, sec, nsec, fr, corner_xyzi, n_corner);
    put_xyzi_cloud(b, sec, nsec, fr, surface_xyzi, n_surface);
    put_xyzi_cloud(b, sec, nsec, fr, nullptr, 0);                                // extracted_normal
  }
  put_blank_cloud(b);                                                            // extracted_ground
  if (out && cap >= b.size()) std::memcpy(out, b.data(), b.size());
  return b.size();
}

/* lidarOdometry.cpp:521-524 — featureLast = corner + surf as (n_corner + n_surface) x 4 floats (x y z intensity).
 * Returns 0, or -1 on a malformed message. xyzi_out may be NULL to query the counts. */
int ro_cloud_info_features(const uint8_t* msg, size_t len, float* xyzi_out, int* n_corner, int* n_surface) {
  Cur c{msg, len, 0, true};
  c.get<uint32_t>(); c.get<uint32_t>(); c.get<uint32_t>(); c.get_string();
  for (int a = 0; a < 4; a++) { const uint32_t k = c.get<uint32_t>(); c.skip(4ull * k); }
  c.skip(36);
  { const uint32_t k = c.get<uint32_t>(); c.skip(4ull * k); }
  c.get<uint8_t>();
  std::vector<float> proj, corner, surf, normal, ground;
  if (!get_cloud(c, proj) || !get_cloud(c, corner) || !get_cloud(c, surf) || !get_cloud(c, normal) || !get_cloud(c, ground) || c.at != len) return -1;
  *n_corner = (int)(corner.size() / 4); *n_surface = (int)(surf.size() / 4);
  if (xyzi_out) {
    if (!corner.empty()) std::memcpy(xyzi_out, corner.data(), corner.size() * 4);
    if (!surf.empty()) std::memcpy(xyzi_out + corner.size(), surf.data(), surf.size() * 4);
  }
  return 0;
}
}
