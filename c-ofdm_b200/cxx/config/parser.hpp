// config/parser.hpp look-alike (reference config/parser.hpp:9-11, parser.cpp:4-33): same ConfigMap type,
// same parse_config signature and semantics ("key = long", '#' comments, lines without '=' skipped,
// std::runtime_error("Cannot open config file"), std::stol on the value, missing keys read 0 via operator[]).
#pragma once
#include <algorithm>
#include <fstream>
#include <stdexcept>
#include <string>
#include <unordered_map>

using ConfigMap = std::unordered_map<std::string, long>;

inline ConfigMap parse_config(const std::string &filename) {
    std::ifstream file(filename);
    if (!file.is_open()) throw std::runtime_error("Cannot open config file");
    ConfigMap cfg;
    std::string line;
    while (std::getline(file, line)) {
        auto notspace = [](unsigned char ch) { return !std::isspace(ch); };
        line.erase(line.begin(), std::find_if(line.begin(), line.end(), notspace));
        line.erase(std::find_if(line.rbegin(), line.rend(), notspace).base(), line.end());
        if (line.empty() || line[0] == '#') continue;
        auto pos = line.find('=');
        if (pos == std::string::npos) continue;
        std::string key = line.substr(0, pos), value = line.substr(pos + 1);
        key.erase(std::remove_if(key.begin(), key.end(), ::isspace), key.end());
        value.erase(std::remove_if(value.begin(), value.end(), ::isspace), value.end());
        cfg[key] = std::stol(value);
    }
    return cfg;
}
