/* tools/Array.hpp — Tools::Array<T>, the container type at the plug-in boundary.
 *
 * Renderer::prerender takes `const Tools::Array<ECS::RenderEntity*>&`
 * (reference src/lib/renderer/Renderer.hpp:48; container declared at reference
 * src/lib/tools/Array.hpp:28-121). This is an independent, std::vector-backed
 * implementation of the members the render path uses, with the reference's
 * names and semantics (resize default-constructs, operator+= appends, wdata /
 * rdata expose the contiguous storage). Inside the reference tree the
 * reference's own header is used instead.
 */
#ifndef RT3_HOST_TOOLS_ARRAY_HPP
#define RT3_HOST_TOOLS_ARRAY_HPP

#include <cstddef>
#include <initializer_list>
#include <limits>
#include <stdexcept>
#include <vector>

namespace Tools {
    template <class T>
    class Array {
        std::vector<T> items;

    public:
        Array() {}
        explicit Array(size_t initial_size) { items.reserve(initial_size); }
        Array(const std::initializer_list<T>& list) : items(list) {}
        Array(const T* list, size_t list_size) : items(list, list + list_size) {}
        Array(const std::vector<T>& list) : items(list) {}

        Array<T>& operator+=(const Array<T>& elems) { items.insert(items.end(), elems.items.begin(), elems.items.end()); return *this; }
        Array<T> operator+(const Array<T>& elems) const { return Array<T>(*this) += elems; }
        void push_back(const T& elem) { items.push_back(elem); }
        void pop_back() { if (!items.empty()) { items.pop_back(); } }
        void erase(size_t index) { if (index < items.size()) { items.erase(items.begin() + (std::ptrdiff_t) index); } }
        void clear() { items.clear(); items.shrink_to_fit(); }
        void reserve(size_t new_size) { if (new_size < items.size()) { items.resize(new_size); } items.reserve(new_size); }
        void resize(size_t new_size) { items.resize(new_size); }

        T& operator[](size_t index) { return items[index]; }
        const T& operator[](size_t index) const { return items[index]; }
        T& at(size_t index) { if (index >= items.size()) { throw std::out_of_range("Index out-of-bounds."); } return items[index]; }
        const T& at(size_t index) const { if (index >= items.size()) { throw std::out_of_range("Index out-of-bounds."); } return items[index]; }

        T* wdata(size_t new_size = std::numeric_limits<size_t>::max()) {
            if (new_size != std::numeric_limits<size_t>::max()) { items.resize(new_size); }
            return items.data();
        }
        const T* rdata() const { return items.data(); }
        bool empty() const { return items.empty(); }
        size_t size() const { return items.size(); }
        size_t capacity() const { return items.capacity(); }
    };
}

#endif
