#ifndef ROLO_REF_SHIM_BOOST_FORMAT
#define ROLO_REF_SHIM_BOOST_FORMAT
#include <ostream>
#include <sstream>
#include <string>
namespace boost {   // only used by the LM debug table (lsq_registration_impl.hpp:114-121): values are appended, not formatted
class format {
  std::string s_;
 public:
  explicit format(const char* f) : s_(f) { s_ += " |"; }
  template <typename T> format& operator%(const T& v) { std::ostringstream os; os << " " << v; s_ += os.str(); return *this; }
  friend std::ostream& operator<<(std::ostream& os, const format& f) { return os << f.s_; }
};
}
#endif
