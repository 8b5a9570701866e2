// ParameterReader.h -- "name = value  # comment" configuration store of the iSS API.
//
// Behavioural contract taken from reference src/ParameterReader.{h,cpp}: names are
// trimmed and lower-cased (ParameterReader.cpp:64-72,130-144), values are doubles,
// later assignments overwrite earlier ones, `getVal(name)` on a missing key prints a
// message and exits with status 1 (ParameterReader.cpp:150-159), the two-argument
// overload returns the default after a message (ParameterReader.cpp:165-175), and
// command-line tokens of the form key=value are accepted (main.cpp:61).
#ifndef ISS_B200_PARAMETERREADER_H_
#define ISS_B200_PARAMETERREADER_H_

#include <string>
#include <utility>
#include <vector>

using std::string;

class ParameterReader {
 public:
    ParameterReader() = default;
    ~ParameterReader() = default;

    void phraseOneLine(string str, string commentSymbol = string("#"));
    void readFromFile(string filename, string commentSymbol = string("#"));
    void readFromArguments(long argc, char *argv[], string commentSymbol = string("#"),
                           long start_from = 1);
    bool exist(string name);
    void setVal(string name, double value);
    double getVal(string name);
    double getVal(string name, double defaultValue);
    // engine addition: like getVal(name, defaultValue) without the "not found" message, for the keys
    // the reference does not know (first_event_index, reduce_checks_over_ranks, nccl_rank, ...)
    double getValQuiet(string name, double defaultValue);
    void echo();

 private:
    std::vector<std::pair<string, double>> entries_;
    long find_(const string &name) const;
    static string normalise_(const string &name);
};

#endif  // ISS_B200_PARAMETERREADER_H_
