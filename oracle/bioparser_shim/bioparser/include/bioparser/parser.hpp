// oracle/bioparser_shim -- TEST INFRASTRUCTURE ONLY.
// A minimal stand-in for the interface of rvaser/bioparser that team_mapper.cpp uses
// (team_mapper.cpp:13-14, :187-188, :230-235, :401-402, :534-551), so that the UNMODIFIED reference
// mapper can be compiled here as an end-to-end oracle. bioparser itself is not vendored by the
// reference (git-ignored, no pinned version) and is not available offline. No hot-path arithmetic
// lives in it: it only turns FASTA/FASTQ text into (name, sequence[, quality]) records.
#ifndef BIOPARSER_SHIM_PARSER_HPP
#define BIOPARSER_SHIM_PARSER_HPP
#include <algorithm>   // the real bioparser headers pull these in; team_mapper.cpp relies on that
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace bioparser {

template <class T>
class Parser {
public:
    virtual ~Parser() = default;
    template <template <class> class P>
    static std::unique_ptr<Parser<T>> Create(const std::string& path) {
        return std::unique_ptr<Parser<T>>(new P<T>(path));
    }
    // Everything is returned by the first call; later calls return an empty vector.
    virtual std::vector<std::unique_ptr<T>> Parse(std::uint64_t bytes, bool shorten_names = true) = 0;

protected:
    explicit Parser(const std::string& path) : in_(path), done_(false) {
        if (!in_) throw std::invalid_argument("[bioparser shim] cannot open " + path);
    }
    static std::string short_name(const std::string& line, bool shorten) {
        std::string n = line.substr(1);
        if (shorten) {
            const auto p = n.find_first_of(" \t");
            if (p != std::string::npos) n.resize(p);
        }
        return n;
    }
    static void chomp(std::string& s) { while (!s.empty() && (s.back() == '\r' || s.back() == '\n')) s.pop_back(); }
    std::ifstream in_;
    bool done_;
};

}  // namespace bioparser
#endif
