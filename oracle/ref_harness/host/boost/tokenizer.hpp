// boost/tokenizer.hpp -- TEST INFRASTRUCTURE (oracle/_ref): the two Boost.Tokenizer types rv/string_utils.cpp uses
// (char_separator with dropped delimiters, no kept delimiters, drop/keep empty tokens), so that the reference file
// compiles where it lies. Boost is not installed in this image. Semantics per the Boost.Tokenizer documentation.
#ifndef SUMA_REF_MINI_BOOST_TOKENIZER
#define SUMA_REF_MINI_BOOST_TOKENIZER
#include <string>
#include <vector>
namespace boost {
enum empty_token_policy { drop_empty_tokens, keep_empty_tokens };
template <class Char>
class char_separator {
 public:
  char_separator(const Char* dropped, const Char* kept = "", empty_token_policy p = drop_empty_tokens)
      : dropped_(dropped ? dropped : ""), kept_(kept ? kept : ""), policy_(p) {}
  std::basic_string<Char> dropped_, kept_;
  empty_token_policy policy_;
};
template <class Sep>
class tokenizer {
 public:
  typedef std::vector<std::string>::const_iterator iterator;
  tokenizer(const std::string& s, const Sep& sep) {
    std::string cur;
    for (size_t i = 0; i <= s.size(); ++i) {
      const bool end = i == s.size();
      const bool is_drop = !end && sep.dropped_.find(s[i]) != std::string::npos;
      const bool is_keep = !end && sep.kept_.find(s[i]) != std::string::npos;
      if (end || is_drop || is_keep) {
        if (!cur.empty() || sep.policy_ == keep_empty_tokens) tokens_.push_back(cur);
        cur.clear();
        if (is_keep) tokens_.push_back(std::string(1, s[i]));
      } else {
        cur.push_back(s[i]);
      }
    }
  }
  iterator begin() const { return tokens_.begin(); }
  iterator end() const { return tokens_.end(); }

 private:
  std::vector<std::string> tokens_;
};
}  // namespace boost
#endif
