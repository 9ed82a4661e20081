// SettingsMap.hpp -- string key/value settings, same interface as the reference's SettingsMap
// (src/Utils/InternalState.hpp:43-126) minus the glm overloads.  ReplayWidget / AutomaticPerformanceMeasurer hand such maps
// to LineRenderer::setNewSettings; the adapter forwards every pair unchanged to lv_set_option.
#pragma once
#include <map>
#include <sstream>
#include <string>

class SettingsMap {
public:
    SettingsMap() = default;
    explicit SettingsMap(const std::map<std::string, std::string>& stringMap) : settings(stringMap) {}
    std::string getValue(const char* key) const { auto it = settings.find(key); return it == settings.end() ? "" : it->second; }
    void addKeyValue(const std::string& key, const std::string& value) { settings[key] = value; }
    template <typename T> void addKeyValue(const std::string& key, const T& value) { std::ostringstream s; s << value; settings[key] = s.str(); }
    void addKeyValue(const std::string& key, bool value) { settings[key] = value ? "true" : "false"; }
    bool isEmpty() const { return settings.empty(); }
    void clear() { settings.clear(); }
    bool getValueOpt(const char* key, std::string& toset) const {
        auto it = settings.find(key);
        if (it == settings.end()) return false;
        toset = it->second;
        return true;
    }
    bool getValueOpt(const char* key, bool& toset) const {  // InternalState.hpp:67-74
        auto it = settings.find(key);
        if (it == settings.end()) return false;
        toset = (it->second == "true") || (it->second == "1");
        return true;
    }
    template <typename T> bool getValueOpt(const char* key, T& toset) const {
        auto it = settings.find(key);
        if (it == settings.end()) return false;
        std::istringstream s(it->second);
        s >> toset;
        return true;
    }
    const std::map<std::string, std::string>& getMap() const { return settings; }

private:
    std::map<std::string, std::string> settings;
};
