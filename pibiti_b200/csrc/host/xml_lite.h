// xml_lite.h -- a small non-validating XML element/attribute reader, enough for Scenes.xml.
// (The reference uses TinyXML, source/SPH/SPH_Scenes.cpp:53-111; only element names, attributes and
// child order matter to the scene loader, so that is all this reader keeps.)
#pragma once
#include <string>
#include <vector>
#include <memory>

namespace sphxml {

struct Element {
    std::string name;
    std::vector<std::pair<std::string, std::string>> attrs;     // in document order
    std::vector<std::unique_ptr<Element>> children;
    Element* parent = nullptr;

    // value of the first attribute with that name, or nullptr (TiXmlElement::Attribute)
    const char* Attribute(const char* key) const;
    // first child element with that name (TiXmlElement::FirstChildElement)
    const Element* FirstChildElement(const char* childName) const;
    // next sibling with that name (TiXmlElement::NextSiblingElement)
    const Element* NextSiblingElement(const char* siblingName) const;
};

struct Document {
    std::unique_ptr<Element> root;      // null if the file could not be read or holds no element
    std::string error;
    bool LoadFile(const char* path);
    bool Parse(const std::string& text);
    const Element* RootElement() const { return root.get(); }
};

}  // namespace sphxml
