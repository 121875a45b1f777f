// boost/lexical_cast.hpp -- TEST INFRASTRUCTURE (oracle/_ref): lexical_cast<arithmetic>(std::string) as documented (the
// whole string must convert, otherwise bad_lexical_cast), so that util/kitti_utils.cpp compiles where it lies.
#ifndef SUMA_REF_MINI_BOOST_LEXICAL_CAST
#define SUMA_REF_MINI_BOOST_LEXICAL_CAST
#include <sstream>
#include <string>
#include <typeinfo>
namespace boost {
class bad_lexical_cast : public std::bad_cast {
 public:
  const char* what() const noexcept override { return "bad lexical cast: source type value could not be interpreted as target"; }
};
template <class Target>
Target lexical_cast(const std::string& s) {
  std::istringstream is(s);
  is.unsetf(std::ios::skipws);
  Target t;
  if (!(is >> t) || is.get() != std::char_traits<char>::eof()) throw bad_lexical_cast();
  return t;
}
}  // namespace boost
#endif
