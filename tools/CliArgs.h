// Minimal stand-in for the `args` library the reference's tools use (Taywee/args, fetched by its CMake): named
// value flags (`--name value`, `--name=value`, `-x value`), boolean flags and positionals, in the reference's spelling.
#ifndef SDFB200_TOOLS_CLI_ARGS_H
#define SDFB200_TOOLS_CLI_ARGS_H

#include <cstdlib>
#include <iostream>
#include <map>
#include <set>
#include <string>
#include <vector>

struct CliArgs
{
    std::map<std::string, std::string> values;
    std::set<std::string> flags;
    std::vector<std::string> positionals;
    bool help = false;

    // valueNames / flagNames: accepted spellings, e.g. {"-d", "--depth"}; aliases maps a short name to its long name
    bool parse(int argc, char** argv, const std::set<std::string>& valueNames, const std::set<std::string>& flagNames,
               const std::map<std::string, std::string>& aliases)
    {
        for (int i = 1; i < argc; i++)
        {
            std::string a = argv[i], v;
            bool hasInline = false;
            const size_t eq = a.find('=');
            if (a.rfind("--", 0) == 0 && eq != std::string::npos) { v = a.substr(eq + 1); a = a.substr(0, eq); hasInline = true; }
            if (aliases.count(a)) a = aliases.at(a);
            if (a == "--help") { help = true; continue; }
            if (flagNames.count(a)) { flags.insert(a); continue; }
            if (valueNames.count(a))
            {
                if (!hasInline) { if (i + 1 >= argc) { std::cerr << "Flag " << a << " needs a value" << std::endl; return false; } v = argv[++i]; }
                values[a] = v;
                continue;
            }
            if (a.rfind("-", 0) == 0 && a.size() > 1 && !(a[1] >= '0' && a[1] <= '9') && a[1] != '.') { std::cerr << "Flag could not be matched: " << a << std::endl; return false; }
            positionals.push_back(a);
        }
        return true;
    }
    bool has(const std::string& n) const { return values.count(n) != 0; }
    std::string str(const std::string& n, const std::string& d) const { return has(n) ? values.at(n) : d; }
    float num(const std::string& n, float d) const { return has(n) ? std::strtof(values.at(n).c_str(), nullptr) : d; }
    uint32_t uint(const std::string& n, uint32_t d) const { return has(n) ? uint32_t(std::strtoul(values.at(n).c_str(), nullptr, 10)) : d; }
};

#endif
