// ParameterReader.cpp -- see ParameterReader.h for the contract and reference citations.
#include "ParameterReader.h"

#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>

namespace {

string strip(const string &s) {
    // the reference's trim() removes every blank and tab, also inner ones (arsenal.cpp:626-640)
    string out;
    out.reserve(s.size());
    for (char c : s)
        if (c != ' ' && c != '\t') out.push_back(c);
    return out;
}

double to_double(const string &s) {
    // stringToDouble of the reference goes through a stringstream extraction
    std::stringstream ss(s);
    double v = 0.0;
    ss >> v;
    return v;
}

}  // namespace

string ParameterReader::normalise_(const string &name) {
    string key = strip(name);
    std::transform(key.begin(), key.end(), key.begin(),
                   [](unsigned char c) { return static_cast<char>(std::tolower(c)); });
    return key;
}

long ParameterReader::find_(const string &name) const {
    const string key = normalise_(name);
    for (size_t i = 0; i < entries_.size(); i++)
        if (entries_[i].first == key) return static_cast<long>(i);
    return -1;
}

void ParameterReader::phraseOneLine(string str, string commentSymbol) {
    if (strip(str).empty()) return;
    const string body = str.substr(0, str.find(commentSymbol));
    if (strip(body).empty()) return;
    const size_t eq = body.find('=');
    if (eq == string::npos) return;     // lines without '=' are silently ignored
    setVal(body.substr(0, eq), to_double(strip(body.substr(eq + 1))));
}

void ParameterReader::readFromFile(string filename, string commentSymbol) {
    std::ifstream in(filename.c_str());
    if (!in) {
        std::cout << "ParameterReader::readFromFile error: file " << filename
                  << " does not exist." << std::endl;
        exit(-1);
    }
    string line;
    while (std::getline(in, line)) phraseOneLine(line, commentSymbol);
}

void ParameterReader::readFromArguments(long argc, char *argv[], string commentSymbol,
                                        long start_from) {
    for (long i = start_from; i < argc; i++) phraseOneLine(argv[i], commentSymbol);
}

bool ParameterReader::exist(string name) { return find_(name) != -1; }

void ParameterReader::setVal(string name, double value) {
    const long idx = find_(name);
    if (idx < 0) {
        entries_.emplace_back(normalise_(name), value);
    } else {
        entries_[idx].second = value;
    }
}

double ParameterReader::getVal(string name) {
    const long idx = find_(name);
    if (idx < 0) {
        std::cout << "ParameterReader::getVal error: parameter with name " << name
                  << " not found." << std::endl;
        exit(1);
    }
    return entries_[idx].second;
}

double ParameterReader::getVal(string name, double defaultValue) {
    const long idx = find_(name);
    if (idx >= 0) return entries_[idx].second;
    std::cout << "ParameterReader::getVal error: parameter with name " << name << " not found."
              << std::endl;
    std::cout << "using defaultValue: " << defaultValue << std::endl;
    return defaultValue;
}

double ParameterReader::getValQuiet(string name, double defaultValue) {
    const long idx = find_(name);
    return idx >= 0 ? entries_[idx].second : defaultValue;
}

void ParameterReader::echo() {
    if (entries_.empty()) return;
    for (auto const &e : entries_) std::cout << e.first << "=" << e.second << "  ";
    std::cout << std::endl;
}
