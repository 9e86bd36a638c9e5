// Facade of ch4/v3/src/funkc.h: the CLI helpers the main loop uses.
#ifndef FUNKC_H
#define FUNKC_H
#include <algorithm>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <utility>
#include <vector>
#include "all.h"

void print_help();
std::string lower(std::string& str);
bool parseArgument(const std::vector<std::string>& args, const std::string& option);

template <class T>
inline void greaterLesser(T& greater, T& lesser) { if (greater < lesser) std::swap(greater, lesser); }

// `--opt value` lookup with a default (funkc.h:28-42)
template <typename T>
T parseArgument(const std::vector<std::string>& args, const std::string& option, T defaultValue) {
    for (size_t i = 0; i + 1 < args.size(); ++i) {
        if (args[i] != option) continue;
        std::istringstream iss(args[i + 1]);
        T value;
        if (iss >> value) { std::cout << " option: " << option << " = " << value << "\n"; return value; }
    }
    std::cout << " option: " << option << " = " << defaultValue << "\n";
    return defaultValue;
}
#endif
