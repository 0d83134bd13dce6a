/**
 * @file IniFile.cpp
 * YARP-style `.ini` text -> IParametersHandler (see IniFile.h for the grammar and its source).
 */
#include <cerrno>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <vector>

#include <BipedalLocomotion/ParametersHandler/IniFile.h>
#include <BipedalLocomotion/ParametersHandler/StdImplementation.h>

namespace
{
enum class Kind
{
    Int,
    Double,
    Bool,
    String
};

struct Token
{
    Kind kind;
    std::string text;
    int i{0};
    double d{0};
    bool b{false};
};

bool isBlank(char c) { return c == ' ' || c == '\t' || c == '\r' || c == ','; }

Token classify(const std::string& word, bool quoted)
{
    Token t{Kind::String, word};
    if (quoted || word.empty()) return t;
    if (word == "true" || word == "false")
    {
        t.kind = Kind::Bool;
        t.b = word == "true";
        return t;
    }
    char* end = nullptr;
    errno = 0;
    const long li = std::strtol(word.c_str(), &end, 10);
    if (*end == '\0' && errno == 0 && li >= -2147483647L - 1 && li <= 2147483647L)
    {
        t.kind = Kind::Int;
        t.i = static_cast<int>(li);
        return t;
    }
    errno = 0;
    const double dv = std::strtod(word.c_str(), &end);
    if (*end == '\0' && end != word.c_str())
    {
        t.kind = Kind::Double;
        t.d = dv;
    }
    return t;
}

/** Split one line into tokens; `inList` is set when a parenthesised list was seen.  Returns false
 * on an unterminated quote or unbalanced parenthesis. */
bool tokenize(const std::string& line, std::vector<Token>& out, bool& sawList)
{
    int depth = 0;
    sawList = false;
    std::size_t p = 0;
    while (p < line.size())
    {
        const char c = line[p];
        if (isBlank(c))
        {
            ++p;
            continue;
        }
        if (c == '#' || (c == '/' && p + 1 < line.size() && line[p + 1] == '/')) break;
        if (c == '(')
        {
            if (++depth > 1) return false; // nested lists are not part of the subset
            sawList = true;
            ++p;
            continue;
        }
        if (c == ')')
        {
            if (--depth < 0) return false;
            ++p;
            continue;
        }
        if (c == '"')
        {
            const std::size_t close = line.find('"', p + 1);
            if (close == std::string::npos) return false;
            out.push_back(classify(line.substr(p + 1, close - p - 1), true));
            p = close + 1;
            continue;
        }
        std::size_t q = p;
        while (q < line.size() && !isBlank(line[q]) && line[q] != '(' && line[q] != ')' && line[q] != '"') ++q;
        out.push_back(classify(line.substr(p, q - p), false));
        p = q;
    }
    return depth == 0;
}

void store(BipedalLocomotion::ParametersHandler::IParametersHandler& h, const std::string& key,
           const std::vector<Token>& v, bool list)
{
    if (!list && v.size() == 1)
    {
        const Token& t = v[0];
        switch (t.kind)
        {
        case Kind::Int: h.setParameter(key, t.i); break;
        case Kind::Double: h.setParameter(key, t.d); break;
        case Kind::Bool: h.setParameter(key, t.b); break;
        case Kind::String: h.setParameter(key, t.text); break;
        }
        return;
    }
    bool allInt = true, allNum = true, allBool = true;
    for (const Token& t : v)
    {
        allInt = allInt && t.kind == Kind::Int;
        allNum = allNum && (t.kind == Kind::Int || t.kind == Kind::Double);
        allBool = allBool && t.kind == Kind::Bool;
    }
    if (v.empty() || allInt)
    {
        std::vector<int> x;
        for (const Token& t : v) x.push_back(t.i);
        h.setParameter(key, x);
    } else if (allNum)
    {
        std::vector<double> x;
        for (const Token& t : v) x.push_back(t.kind == Kind::Int ? static_cast<double>(t.i) : t.d);
        h.setParameter(key, x);
    } else if (allBool)
    {
        std::vector<bool> x;
        for (const Token& t : v) x.push_back(t.b);
        h.setParameter(key, x);
    } else
    {
        std::vector<std::string> x;
        for (const Token& t : v) x.push_back(t.text);
        h.setParameter(key, x);
    }
}
} // namespace

namespace BipedalLocomotion
{
namespace ParametersHandler
{

bool loadIniString(const std::string& text, IParametersHandler& handler)
{
    std::istringstream in(text);
    std::string line;
    IParametersHandler* current = &handler;
    std::shared_ptr<StdImplementation> group;
    int number = 0;
    while (std::getline(in, line))
    {
        ++number;
        std::size_t first = 0;
        while (first < line.size() && isBlank(line[first])) ++first;
        if (first == line.size()) continue;
        if (line[first] == '[')
        {
            const std::size_t close = line.find(']', first);
            std::string name = close == std::string::npos ? "" : line.substr(first + 1, close - first - 1);
            while (!name.empty() && isBlank(name.back())) name.pop_back();
            while (!name.empty() && isBlank(name.front())) name.erase(name.begin());
            if (name.empty())
            {
                std::cerr << "[loadIniString] Malformed group header at line " << number << "." << std::endl;
                return false;
            }
            group = std::make_shared<StdImplementation>();
            if (!handler.setGroup(name, group))
            {
                std::cerr << "[loadIniString] Unable to create the group " << name << " (line " << number
                          << ")." << std::endl;
                return false;
            }
            current = group.get();
            continue;
        }
        std::vector<Token> tokens;
        bool sawList = false;
        if (!tokenize(line, tokens, sawList))
        {
            std::cerr << "[loadIniString] Unterminated quote or unbalanced parenthesis at line " << number
                      << "." << std::endl;
            return false;
        }
        if (tokens.empty()) continue; // comment-only line
        if (tokens.size() == 1 && !sawList)
        {
            std::cerr << "[loadIniString] The key " << tokens[0].text << " has no value (line " << number
                      << ")." << std::endl;
            return false;
        }
        const std::string key = tokens[0].text;
        tokens.erase(tokens.begin());
        store(*current, key, tokens, sawList || tokens.size() > 1);
    }
    return true;
}

bool loadIniFile(const std::string& path, IParametersHandler& handler)
{
    std::ifstream f(path);
    if (!f)
    {
        std::cerr << "[loadIniFile] Unable to open " << path << "." << std::endl;
        return false;
    }
    std::stringstream ss;
    ss << f.rdbuf();
    return loadIniString(ss.str(), handler);
}

} // namespace ParametersHandler
} // namespace BipedalLocomotion
