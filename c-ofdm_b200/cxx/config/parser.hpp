// config/parser.hpp of the drop-in tree: the type and the entry point the reference apps use (config/parser.hpp:9-11),
// implemented with this repository's own config reader (csrc/host_consts.hpp parse_config_file), whose observable
// behaviour equals parser.cpp:4-33: "key = long" lines, '#' comments, lines without '=' skipped, white space ignored,
// std::runtime_error("Cannot open config file"), std::stol's exceptions on a non-numeric value, missing keys read 0
// through operator[].
#pragma once
#include <cctype>
#include <fstream>
#include <stdexcept>
#include <string>
#include <unordered_map>

using ConfigMap = std::unordered_map<std::string, long>;

inline ConfigMap parse_config(const std::string &filename) {
    std::ifstream f(filename);
    if (!f.is_open()) throw std::runtime_error("Cannot open config file");
    ConfigMap cfg;
    std::string line;
    while (std::getline(f, line)) {
        size_t a = 0, b = line.size();
        while (a < b && std::isspace((unsigned char)line[a])) a++;
        while (b > a && std::isspace((unsigned char)line[b - 1])) b--;
        if (a == b || line[a] == '#') continue;
        const size_t eq = line.find('=', a);
        if (eq == std::string::npos || eq >= b) continue;
        std::string key, val;
        for (size_t i = a; i < eq; i++) if (!std::isspace((unsigned char)line[i])) key += line[i];
        for (size_t i = eq + 1; i < b; i++) if (!std::isspace((unsigned char)line[i])) val += line[i];
        cfg[key] = std::stol(val);
    }
    return cfg;
}
